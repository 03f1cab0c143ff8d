# differential timing of the B=200 schedule: which launches are on the critical path? (results are wrong with a mask set)
for m in ${MASKS:-0 1 2 3 4 7 8 16 24 32 64 96 127}; do
CLV_DIAG_SKIP=$m timeout 60 python bench.py --no-sampler --no-vae --no-cpu --steps 300 2>/dev/null | python -c "import sys,json
d=json.loads(sys.stdin.readlines()[-1]); print('skip=%3d  %.2f us/step  %d launches' % ($m, d['ms_per_step']*1e3, d['launches_per_step']))"
done

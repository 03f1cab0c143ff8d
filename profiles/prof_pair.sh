#!/bin/bash
# ncu stall sampling of the pair kernels on the micro-benchmark (run on the GPU box through gpurun)
O=gpurun_out/pp; mkdir -p $O
K=${1:-lstm_pair_fwd_kernel}
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -o $O/prof_$K -f python profiles/pbench.py 200 16 > $O/ncu_$K.log 2>&1
ncu -i $O/prof_$K.ncu-rep --page source --csv > $O/src_$K.csv 2>/dev/null
python profiles/stalls.py $O/src_$K.csv 45 > $O/stalls_$K.txt
python profiles/rawsum.py $O/prof_$K.ncu-rep > $O/raw_$K.md 2>/dev/null
rm -f $O/src_$K.csv $O/prof_$K.ncu-rep
cat $O/stalls_$K.txt

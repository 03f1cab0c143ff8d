"""
Device-side engine behind the Keras-like model objects: owns the flat parameter / gradient /
optimizer-state / workspace buffers (PyTorch is only the allocator and the stream/NCCL plumbing) and
drives libclv_b200's train step -- on one GPU with the Adam-WN update scheduled inside it
(clv_train_step_opt), with N > 1 as step -> NCCL gradient all-reduce -> Adam-WN -- optionally replayed
as one CUDA graph.  One process per GPU; data parallel = batch sharded across ranks.

Replaces what Keras' `train_function` / `test_function` do for the reference (cl_vrnn/train.py:66-71;
graph cl_vrnn/model.py:169-264; optimizer utils/weightnorm.py:75-143).
"""
import ctypes as C
import math
import os
import numpy as np
import torch

from . import _lib
from ._lib import lib, check, ptr, clv_adam_args, clv_p2p_args

VRNN_TENSORS = ["hW.kernel", "hW.bias", "Wargs.kernel", "Wargs.bias",
                "encoder_h.kernel", "encoder_h.recurrent_kernel", "encoder_h.bias",
                "Z_mean.kernel", "Z_mean.bias", "Z_log_var.kernel", "Z_log_var.bias",
                "decoder_h.kernel", "decoder_h.recurrent_kernel", "decoder_h.bias",
                "X_decoded_mean.kernel", "X_decoded_mean.bias"]
VAE_TENSORS = ["h_w.kernel", "h_w.bias", "w_mean.kernel", "w_mean.bias", "w_log_var.kernel",
               "w_log_var.bias", "h.kernel", "h.bias", "z_mean.kernel", "z_mean.bias",
               "z_log_var.kernel", "z_log_var.bias", "decoder_h.kernel", "decoder_h.bias",
               "x_decoded_mean.kernel", "x_decoded_mean.bias"]
LOSS_NAMES = ["vae", "w_kl", "w_rec", "z_kl", "acc"]


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda():
    if not torch.cuda.is_available():
        raise _lib.ClvError("no CUDA device: the CL-VAE/CL-VRNN hot path has no CPU fallback")


class Engine:
    def __init__(self, model, B, L=1, D=88, H=88, Z=2, n_classes=2, use_x_prev=False, Hc=88,
                 class_weight=1.0, kl_weight=1.0, w_kl_weight=1.0, w_log_var_prior=0.0,
                 optimizer="adam-wn", lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-8,
                 seed=0, device=None, world_size=1, rank=0, process_group=None, use_graph=True,
                 overlap_wgrad=True, gemm_algo=1, p2p_allreduce=None, tc_lstm_min=0,
                 fused_optimizer=True, predict_next=False, micro_batch=None, workspace_budget_bytes=None,
                 pair_bwd=False):
        _require_cuda()
        lib()
        if optimizer not in ("adam-wn", "adam"):
            raise NotImplementedError("optimizer %r: only 'adam-wn' (reference default) and 'adam' "
                                      "are built" % (optimizer,))
        self.model = {"vrnn": 0, "vae": 1}[model] if isinstance(model, str) else int(model)
        self.dev = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.B, self.L, self.D, self.H, self.Z, self.C = B, (L if self.model == 0 else 1), D, H, Z, n_classes
        self.Hc = Hc
        self.use_x_prev = bool(use_x_prev)
        self.pair_bwd = bool(pair_bwd)                                 # both BPTTs as one wavefront launch (opt-in)
        self.predict_next = bool(predict_next)                         # --predict_next: target = next frame
        if self.predict_next and self.use_x_prev:
            raise ValueError("Can't use --predict_next if using --use_x_prev")     # cl_vrnn/train.py:28
        self.W = self.L + 1 if (self.use_x_prev or self.predict_next) else self.L   # frames per window
        self.x_shift = 0                                               # 0 = default placement
        self.y_shift = 1 if self.predict_next else 0                   # 0 = the target is `current`
        self.hyper = dict(class_weight=class_weight, kl_weight=kl_weight, w_kl_weight=w_kl_weight,
                          w_log_var_prior=w_log_var_prior)
        self.optimizer, self.lr, self.b1, self.b2, self.eps = optimizer, lr, beta_1, beta_2, epsilon
        self.world_size, self.rank, self.pg = world_size, rank, process_group
        self.seed = (int(seed) * 1000003 + rank * 7919 + 1) & 0xFFFFFFFFFFFFFFFF
        self.use_graph = use_graph
        self.overlap_wgrad = bool(overlap_wgrad)
        self.tc_lstm_min = int(tc_lstm_min)        # 0 = library default (16384)
        self.fused_optimizer = bool(fused_optimizer)   # world_size 1: clv_train_step_opt
        self.gemm_algo = int(gemm_algo)            # 0: exact-fp32 SIMT GEMMs, 1: tcgen05 input projections
        if self.overlap_wgrad:
            with torch.cuda.device(self.dev):
                check(lib().clv_runtime_init(), "clv_runtime_init")
        self.names = VRNN_TENSORS if self.model == 0 else VAE_TENSORS
        # host micro-batching: a logical batch of B sequences is stepped as B / Bm calls of Bm sequences that
        # ACCUMULATE into the same gradient / loss buffers (clv_cfg.accumulate), then one exchange + Adam-WN.
        # The workspace (activations of one call) is sized for Bm, so the top of the sweep (B = 65 536, L = 512:
        # 177 GB of activations) fits.  micro_batch=None: the largest divisor of B whose workspace fits the budget
        # (default 60 % of the free HBM).
        self.Bm = B
        if micro_batch is not None:
            if B % int(micro_batch):
                raise ValueError("micro_batch %d must divide the batch %d" % (micro_batch, B))
            self.Bm = int(micro_batch)
        else:
            budget = workspace_budget_bytes
            if budget is None:
                free, _ = torch.cuda.mem_get_info(self.dev)
                budget = int(0.6 * free)
            while self.Bm > 1 and lib().clv_workspace_bytes(C.byref(self.cfg(B=self.Bm))) > budget:
                nb = self.Bm - 1
                while nb > 1 and B % nb:
                    nb -= 1
                self.Bm = nb
        self.n_micro = B // self.Bm
        cfg = self.cfg()
        self.P, self.offs, self.rows, self.cols = _lib.param_layout(cfg)
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.params = torch.zeros(self.P, **f32)
        # [grads | 8 loss scalars].  Data parallel (world_size > 1), all inside clv_train_step_opt and the step's
        # CUDA graph:
        #  * p2p_allreduce=None (default: used when symmetric memory can be set up, the optimizer is scheduled
        #    inside the step and the batch is not micro-batched) / True: the buffer and a small flag block live in
        #    peer-mapped (symmetric) memory; per gradient bucket a one-shot all-reduce KERNEL publishes the bucket
        #    (flag store into every peer's flag block), waits for the peers' flags inside the kernel and reads
        #    their gradients over NVLink in rank order; the ordinary Adam-WN update of the bucket follows
        #    (clv_p2p_allreduce).  Buckets: [decoder | X head | losses] hidden behind the encoder BPTT, [key
        #    encoder] and [encoder LSTM | Z heads] concurrently on the tail.  N=2: 0.158 ms/step, N=8: 0.177
        #    (end to end 0.199);
        #  * p2p_allreduce="fused": the exchange inside the Adam-WN kernels themselves (clv_adamwn_step_range_p2p;
        #    strided 16..32-byte remote reads: slower);
        #  * p2p_allreduce=False: ONE NCCL all-reduce of the whole buffer, enqueued by the library through the
        #    exchange callback after the last weight gradient, then one Adam-WN launch (N=2: 0.161-0.163 ms/step,
        #    N=8: 0.179, end to end 0.264; two uncoupled ranks: 0.139).
        if p2p_allreduce is None and (world_size == 1 or not fused_optimizer or self.n_micro > 1):
            p2p_allreduce = False
        self.symm = None
        self.p2p = None
        if world_size == 1 and p2p_allreduce is not False and fused_optimizer:
            # diagnostic: the peer-memory schedule with this rank as its only peer (what the schedule itself
            # costs, without NVLink and without another rank to wait for)
            self.gradbuf = torch.zeros(self.P + 8, **f32)
            self.peer_ptrs = torch.tensor([self.gradbuf.data_ptr()], dtype=torch.int64, device=self.dev)
            self.flagbuf = torch.zeros(int(lib().clv_p2p_flag_ints()), dtype=torch.int32, device=self.dev)
            self.peer_flag_ptrs = torch.tensor([self.flagbuf.data_ptr()], dtype=torch.int64, device=self.dev)
            self.gsum = torch.zeros(self.P + 8, **f32)
            self.loss_red = self.gsum[self.P:]
            self.p2p = clv_p2p_args(peer_grads=self.peer_ptrs.data_ptr(), peer_flags=self.peer_flag_ptrs.data_ptr(),
                                    n_peers=1, rank=0, gsum=self.gsum.data_ptr(), loss_out=self.loss_red.data_ptr(),
                                    form=1 if p2p_allreduce == "fused" else 0)
        if world_size > 1 and p2p_allreduce is not False:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                group = process_group if process_group is not None else torch.distributed.group.WORLD
                self.gradbuf = symm_mem.empty(self.P + 8, dtype=torch.float32, device=self.dev)
                self.gradbuf.zero_()
                self.symm = symm_mem.rendezvous(self.gradbuf, group)
                self.peer_ptrs = torch.tensor([int(x) for x in self.symm.buffer_ptrs], dtype=torch.int64,
                                              device=self.dev)
                nfl = int(lib().clv_p2p_flag_ints())
                self.flagbuf = symm_mem.empty(nfl, dtype=torch.int32, device=self.dev)
                self.flagbuf.zero_()
                self.symm_flags = symm_mem.rendezvous(self.flagbuf, group)
                self.peer_flag_ptrs = torch.tensor([int(x) for x in self.symm_flags.buffer_ptrs], dtype=torch.int64,
                                                   device=self.dev)
                # the reduced [grads | losses].  CLV_P2P_MC=1 (opt-in) sends the exchange through the NVSwitch
                # multicast mapping of the two buffers instead (two-shot: multimem.ld_reduce of this rank's 1/N
                # slice + multimem.st of the sum to every rank; gsum is symmetric too then).  Measured equal at
                # N=8 (0.178 vs 0.177 ms/step: 7-9 us per bucket against 11-13) and slower at N=2 (0.176 vs 0.158)
                want_mc = os.environ.get("CLV_P2P_MC") == "1" and p2p_allreduce != "fused"
                mc_grads = int(getattr(self.symm, "multicast_ptr", 0) or 0) if want_mc else 0
                mc_gsum = 0
                if mc_grads:
                    self.gsum = symm_mem.empty(self.P + 8, dtype=torch.float32, device=self.dev)
                    self.gsum.zero_()
                    self.symm_gsum = symm_mem.rendezvous(self.gsum, group)
                    mc_gsum = int(getattr(self.symm_gsum, "multicast_ptr", 0) or 0)
                else:
                    self.gsum = torch.zeros(self.P + 8, **f32)
                if not mc_gsum:
                    mc_grads = 0
                self.loss_red = self.gsum[self.P:]
                self.p2p = clv_p2p_args(peer_grads=self.peer_ptrs.data_ptr(), peer_flags=self.peer_flag_ptrs.data_ptr(),
                                        n_peers=world_size, rank=rank, gsum=self.gsum.data_ptr(),
                                        loss_out=self.loss_red.data_ptr(),
                                        form=1 if p2p_allreduce == "fused" else 0,
                                        mc_grads=mc_grads or None, mc_gsum=mc_gsum or None)
                if os.environ.get("CLV_P2P_DIAG_SELF") == "1":
                    # diagnostic (WRONG results): symmetric buffers, but every rank exchanges with itself only
                    self.peer_ptrs = self.peer_ptrs[rank:rank + 1].clone()
                    self.peer_flag_ptrs = self.peer_flag_ptrs[rank:rank + 1].clone()
                    self.p2p.peer_grads, self.p2p.peer_flags = self.peer_ptrs.data_ptr(), self.peer_flag_ptrs.data_ptr()
                    self.p2p.n_peers, self.p2p.rank = 1, 0
                torch.cuda.synchronize(self.dev)
                torch.distributed.barrier(group=process_group)        # every rank's flags are zero before any signal
            except Exception as ex:  # noqa: BLE001
                if p2p_allreduce is not None:
                    raise
                print("[clvae_b200] symmetric memory unavailable (%s): using the NCCL all-reduce path" % (ex,))
                self.symm = None
                self.p2p = None
            if p2p_allreduce is None:
                # automatic mode: every rank must have taken the same decision
                ok = torch.tensor([1 if self.p2p is not None else 0], dtype=torch.int32, device=self.dev)
                torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN, group=process_group)
                if int(ok.item()) == 0:
                    self.symm = None
                    self.p2p = None
        if self.symm is None and self.p2p is None:
            self.gradbuf = torch.zeros(self.P + 8, **f32)
        self.grads, self.loss_acc = self.gradbuf[:self.P], self.gradbuf[self.P:]
        self._loss_src = self.loss_acc
        n_state = check(lib().clv_adamwn_state_floats(C.byref(cfg)), "clv_adamwn_state_floats")
        self.opt_state = torch.zeros(n_state, **f32)
        check(lib().clv_adamwn_init(C.byref(cfg), ptr(self.opt_state), _stream()), "clv_adamwn_init")
        ws_bytes = check(lib().clv_workspace_bytes(C.byref(self.cfg(B=self.Bm))), "clv_workspace_bytes")
        self.workspace = torch.zeros(ws_bytes // 4 + 64, **f32)
        self.eps_w = torch.zeros(B * (self.C - 1), **f32)
        self.eps_z = torch.zeros(B * self.L * Z, **f32)
        self.rng_ctr = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self._win_off_seq = (torch.arange(B, dtype=torch.int32, device=self.dev) * self.W).contiguous()
        self.win_off = self._win_off_seq.clone()
        self._win_off_is_seq = True
        self._graphs = {}
        self._alloc_window_staging()
        self.roll = self.win_buf                                       # or a resident dataset roll
        self.loss_host = torch.zeros(8, dtype=torch.float32).pin_memory()
        self._loss_mirrored = False      # loss_host already holds the last step's scalars
        self.launches_per_step = 0
        self._exchange_cb = _lib.EXCHANGE_FN(self._exchange)     # kept alive: the library holds a raw pointer
        self._exchange_error = None

    def _alloc_window_staging(self):
        """One device allocation [windows uint8 | pad | labels int32]: a host batch whose labels sit
        right behind its windows in ONE pinned buffer is staged with a single H2D copy."""
        nw = self.B * self.W * self.D
        self._win_pad = (-nw) % 4
        self._winlab = torch.zeros(nw + self._win_pad + 4 * self.B, dtype=torch.uint8, device=self.dev)
        self.win_buf = self._winlab[:nw]
        self.labels = self._winlab[nw + self._win_pad:].view(torch.int32)
        self._graphs.clear()

    # ------------------------------------------------------------------ configuration
    def cfg(self, **over):
        kw = dict(model=self.model, B=self.B, L=self.L, D=self.D, H=self.H, Z=self.Z, C_=self.C,
                  use_x_prev=self.use_x_prev, Hc=self.Hc, B_global=self.B * self.world_size,
                  seed=self.seed, x_shift=self.x_shift, y_shift=self.y_shift, overlap_wgrad=int(self.overlap_wgrad), gemm_algo=self.gemm_algo,
                  tc_lstm_min=self.tc_lstm_min, pair_bwd=int(self.pair_bwd),
                  **self.hyper)
        kw.update(over)
        return _lib.make_cfg(**kw)

    def set_loss_weights(self, **kw):
        """kl_weight / w_kl_weight annealing (utils/model_utils.py:41-49): graphs are re-captured."""
        changed = any(self.hyper.get(k) != v for k, v in kw.items())
        self.hyper.update(kw)
        if changed:
            self._graphs.clear()

    def set_window(self, frames, x_shift, y_shift=None):
        """Window geometry: `frames` per window, `current` starts at frame x_shift (0 = default:
        1 with use_x_prev), the target at y_shift (0 = the target is `current`).  Used when
        current/history (or input/target) are independent arrays ([history | current], [x | y])."""
        self.W, self.x_shift = frames, x_shift
        if y_shift is not None:
            self.y_shift = y_shift
        self._win_off_seq = (torch.arange(self.B, dtype=torch.int32, device=self.dev) * self.W).contiguous()
        self.win_off.copy_(self._win_off_seq)
        self._win_off_is_seq = True
        self._alloc_window_staging()
        self.roll = self.win_buf
        self._graphs.clear()

    # ------------------------------------------------------------------ parameters
    def view(self, name):
        i = self.names.index(name)
        n = (self.rows[i] if self.rows[i] > 0 else 1) * self.cols[i]
        t = self.params[self.offs[i]:self.offs[i] + n]
        return t.view(self.rows[i], self.cols[i]) if self.rows[i] > 0 else t

    def grad_view(self, name):
        i = self.names.index(name)
        n = (self.rows[i] if self.rows[i] > 0 else 1) * self.cols[i]
        t = self.grads[self.offs[i]:self.offs[i] + n]
        return t.view(self.rows[i], self.cols[i]) if self.rows[i] > 0 else t

    def set_params(self, d):
        for k in self.names:
            self.view(k).copy_(torch.as_tensor(np.asarray(d[k]), dtype=torch.float32).to(self.dev))

    def get_params(self):
        return {k: self.view(k).detach().cpu().numpy().copy() for k in self.names}

    def init_params(self, rng):
        """Keras-2.0.0 default initialisers [K2-recall]: Dense glorot_uniform / zeros; LSTM kernel
        glorot_uniform, recurrent orthogonal, unit_forget_bias; VRNN Z heads and X head
        RandomNormal(0, 0.1) (cl_vrnn/model.py:200-207,229-233)."""
        out = {}
        for i, k in enumerate(self.names):
            r, c = self.rows[i], self.cols[i]
            layer = k.split(".")[0]
            if r == 0:
                v = np.zeros(c, np.float32)
                if self.model == 0 and layer in ("encoder_h", "decoder_h"):
                    v[self.H:2 * self.H] = 1.0
            elif k.endswith("recurrent_kernel"):
                a = rng.standard_normal((r, c))
                u, _, vt = np.linalg.svd(a, full_matrices=False)
                v = (u if u.shape == (r, c) else vt).astype(np.float32)
            elif self.model == 0 and layer in ("Z_mean", "Z_log_var", "X_decoded_mean"):
                v = rng.normal(0.0, 0.1, (r, c)).astype(np.float32)
            else:
                lim = math.sqrt(6.0 / (r + c))
                v = rng.uniform(-lim, lim, (r, c)).astype(np.float32)
            out[k] = v
        self.set_params(out)
        if self.world_size > 1:
            torch.distributed.broadcast(self.params, src=0, group=self.pg)
        return out

    # ------------------------------------------------------------------ data staging
    def set_resident_roll(self, roll_u8):
        """Keep a whole split on the device: uint8 [n_frames, D]; batches are then just window
        offsets (utils.pianoroll.DeviceRolls)."""
        self.roll = torch.as_tensor(roll_u8, dtype=torch.uint8).to(self.dev).contiguous().view(-1)

    def stage_windows(self, win_u8, labels_i32, non_blocking=True):
        """H2D of one batch given as materialised windows [B, W, D] uint8 + int32 labels."""
        self.roll = self.win_buf
        if not self._win_off_is_seq:                                   # window b = frames [b*W, (b+1)*W)
            self.win_off.copy_(self._win_off_seq)
            self._win_off_is_seq = True
        w = win_u8.reshape(-1)
        nw = w.numel()
        if (self._win_pad == 0 and not w.is_cuda and not labels_i32.is_cuda and labels_i32.dtype == torch.int32
                and labels_i32.is_contiguous() and w.is_contiguous()
                and labels_i32.data_ptr() == w.data_ptr() + nw
                and labels_i32.untyped_storage().data_ptr() == w.untyped_storage().data_ptr()):
            # labels packed behind the windows in one host buffer: one copy for both
            self._winlab.copy_(w.as_strided((nw + 4 * self.B,), (1,)), non_blocking=non_blocking)
        else:
            self.win_buf.copy_(w, non_blocking=non_blocking)
            self.labels.copy_(labels_i32, non_blocking=non_blocking)

    def stage_offsets(self, off_i32, labels_i32, non_blocking=True):
        self._win_off_is_seq = False
        self.win_off.copy_(off_i32, non_blocking=non_blocking)
        self.labels.copy_(labels_i32, non_blocking=non_blocking)

    # ------------------------------------------------------------------ one step
    def _launch_step(self, train, gen_noise):
        cfg = self.cfg(gen_noise=int(gen_noise), do_backward=int(train))
        if self.n_micro > 1:
            return self._launch_micro(train, gen_noise)
        if train and self.fused_optimizer:
            # the optimizer is part of the step's schedule (Adam-WN per tensor range, overlapped with the
            # encoder BPTT / wgrads).  Data parallel: either the peer-memory form (self.p2p: signal kernels +
            # Adam-WN kernels that sum the peers' gradients over NVLink) or the NCCL callback (one all-reduce)
            opt = clv_adam_args(state=self.opt_state.data_ptr(), lr=self.lr, beta_1=self.b1, beta_2=self.b2,
                                epsilon=self.eps, grad_scale=1.0, weightnorm=int(self.optimizer == "adam-wn"),
                                loss_mirror=self.loss_host.data_ptr(),   # pinned => device-visible (UVA)
                                exchange=(C.cast(self._exchange_cb, C.c_void_p)
                                          if (self.world_size > 1 and self.p2p is None) else None),
                                exchange_user=None,
                                p2p=(C.addressof(self.p2p) if self.p2p is not None else None))
            self._exchange_error = None
            rc = lib().clv_train_step_opt(C.byref(cfg), ptr(self.params), ptr(self.grads), ptr(self.loss_acc),
                                          ptr(self.roll), ptr(self.win_off), ptr(self.labels),
                                          ptr(self.eps_w), ptr(self.eps_z), ptr(self.rng_ctr),
                                          ptr(self.workspace), self.workspace.numel() * 4, C.byref(opt),
                                          _stream())
            if self._exchange_error is not None:
                raise _lib.ClvError("gradient exchange callback failed") from self._exchange_error
            check(rc, "clv_train_step_opt")
            return
        check(lib().clv_train_step(C.byref(cfg), ptr(self.params), ptr(self.grads), ptr(self.loss_acc),
                                   ptr(self.roll), ptr(self.win_off), ptr(self.labels),
                                   ptr(self.eps_w), ptr(self.eps_z), ptr(self.rng_ctr),
                                   ptr(self.workspace), self.workspace.numel() * 4, _stream()),
              "clv_train_step")
        if train and self.symm is not None:
            # fused path: barrier (all ranks' gradients written) -> Adam-WN reads the peers' buffers over
            # NVLink and reduces in-kernel -> barrier (all peers done reading before anyone overwrites)
            self.symm.barrier(channel=0)
            check(lib().clv_adamwn_step_p2p(C.byref(cfg), ptr(self.params), ptr(self.peer_ptrs),
                                            self.world_size, ptr(self.gsum), ptr(self.loss_red),
                                            ptr(self.opt_state), self.lr, self.b1, self.b2, self.eps,
                                            int(self.optimizer == "adam-wn"), _stream()), "clv_adamwn_step_p2p")
            self.symm.barrier(channel=1)
            return
        if self.world_size > 1:
            buf = self.gradbuf if train else self.loss_acc
            torch.distributed.all_reduce(buf, group=self.pg)
        if train:
            check(lib().clv_adamwn_step(C.byref(cfg), ptr(self.params), ptr(self.grads),
                                        ptr(self.opt_state), self.lr, self.b1, self.b2, self.eps, 1.0,
                                        int(self.optimizer == "adam-wn"), _stream()), "clv_adamwn_step")

    def _launch_micro(self, train, gen_noise):
        """B / Bm accumulated calls of Bm sequences, then the exchange and ONE Adam-WN update (SURVEY 7 "BPTT
        memory at the top of the sweep"; test_microbatch_accumulation_equals_full_batch)."""
        if self.symm is not None:
            raise NotImplementedError("micro-batching with the peer-memory exchange")
        Bm, C1 = self.Bm, self.C - 1
        for k in range(self.n_micro):
            cfg = self.cfg(B=Bm, gen_noise=int(gen_noise), do_backward=int(train), accumulate=int(k > 0))
            check(lib().clv_train_step(C.byref(cfg), ptr(self.params), ptr(self.grads), ptr(self.loss_acc),
                                       ptr(self.roll), ptr(self.win_off[k * Bm:]), ptr(self.labels[k * Bm:]),
                                       ptr(self.eps_w[k * Bm * C1:]), ptr(self.eps_z[k * Bm * self.L * self.Z:]),
                                       ptr(self.rng_ctr), ptr(self.workspace), self.workspace.numel() * 4, _stream()),
                  "clv_train_step")
        if self.world_size > 1:
            torch.distributed.all_reduce(self.gradbuf if train else self.loss_acc, group=self.pg)
        if train:
            check(lib().clv_adamwn_step(C.byref(self.cfg()), ptr(self.params), ptr(self.grads),
                                        ptr(self.opt_state), self.lr, self.b1, self.b2, self.eps, 1.0,
                                        int(self.optimizer == "adam-wn"), _stream()), "clv_adamwn_step")

    def _exchange(self, user, buf, count, stream):
        """clv_exchange_fn: sum-all-reduce gradbuf[buf .. buf+count) over the ranks on `stream`."""
        try:
            off = (int(buf) - self.gradbuf.data_ptr()) // 4
            view = self.gradbuf[off:off + int(count)]
            sp = int(stream or 0)              # a NULL cudaStream_t (the legacy default stream) arrives as None
            ctx = torch.cuda.stream(torch.cuda.ExternalStream(sp, device=self.dev)) if sp else torch.cuda.stream(torch.cuda.default_stream(self.dev))
            with ctx:
                torch.distributed.all_reduce(view, group=self.pg)
            return 0
        except Exception as ex:  # noqa: BLE001  (an exception must not cross the C frame)
            self._exchange_error = ex
            return -1

    def run(self, train=True, gen_noise=True):
        """Launch one step on the current stream using the staged batch.  With use_graph the launch
        sequence (fwd+bwd kernels, NCCL all-reduce, Adam-WN) is captured once per
        (train, gen_noise, roll buffer) and replayed."""
        # where the (globally reduced) loss scalars of this step end up
        self._loss_src = self.loss_red if (train and self.symm is not None) else self.loss_acc
        # the scheduled-optimizer step mirrors its loss scalars into loss_host from its last kernel
        self._loss_mirrored = bool(train and self.fused_optimizer and self.n_micro == 1)
        if not self.use_graph:
            n0 = lib().clv_launch_count()
            self._launch_step(train, gen_noise)
            if train:
                self.launches_per_step = lib().clv_launch_count() - n0
            return
        key = (bool(train), bool(gen_noise), self.roll.data_ptr())
        g = self._graphs.get(key)
        if g is None:
            s = torch.cuda.Stream(device=self.dev)
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                g = torch.cuda.CUDAGraph()
                n0 = lib().clv_launch_count()
                with torch.cuda.graph(g, stream=s):
                    self._launch_step(train, gen_noise)
                if train:
                    self.launches_per_step = lib().clv_launch_count() - n0
            torch.cuda.current_stream().wait_stream(s)
            self._graphs[key] = g
            # capture does not execute: fall through to replay
        g.replay()

    def p2p_stats(self, reset=True):
        """Diagnostics of the peer-memory exchange: per gradient bucket, the mean time (us) block 0 of the
        all-reduce kernel spent waiting for the peers' flags and moving its share of the data."""
        if self.p2p is None:
            return None
        torch.cuda.synchronize(self.dev)
        row = self.flagbuf[self.flagbuf.numel() - 16:]      # the local row of the flag block
        v = row.cpu().tolist()
        out = {}
        for slot in range(4):
            w, c, k = v[1 + 3 * slot:4 + 3 * slot]
            if k:
                out["bucket%d" % slot] = {"wait_us": round(w / k / 1e3, 2), "copy_us": round(c / k / 1e3, 2), "n": k}
        if reset:
            row[1:13].zero_()
        return out

    def read_losses(self):
        """D2H of the 5 scalars (already global means) -> dict incl. Keras' weighted total."""
        if not self._loss_mirrored:
            self.loss_host.copy_(self._loss_src, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        v = self.loss_host.tolist()
        d = dict(zip(LOSS_NAMES, v[:5]))
        d["loss"] = (d["vae"] + self.hyper["w_kl_weight"] * d["w_kl"]
                     + self.hyper["class_weight"] * d["w_rec"] + self.hyper["kl_weight"] * d["z_kl"])
        return d

    def ws_view(self, name, shape):
        off = lib().clv_workspace_offset(C.byref(self.cfg()), name.encode())
        if off < 0:
            raise KeyError(name)
        n = int(np.prod(shape))
        return self.workspace[off:off + n].view(*shape)

    @property
    def iterations(self):
        return int(self.opt_state[-2:-1].view(torch.int32).item())

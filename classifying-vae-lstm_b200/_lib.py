"""
ctypes binding of libclv_b200.so (include/clv_b200.h).  There is NO fallback: if the shared library
is missing or an entry point returns an error, an exception is raised -- nothing is ever routed to
PyTorch ops or the CPU oracle.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libclv_b200.so")

CLV_N_TENSORS = 16


class ClvError(RuntimeError):
    pass


class clv_cfg(C.Structure):
    _fields_ = [("model", C.c_int32), ("B", C.c_int32), ("B_global", C.c_int32), ("L", C.c_int32),
                ("D", C.c_int32), ("H", C.c_int32), ("Hc", C.c_int32), ("Z", C.c_int32),
                ("C", C.c_int32), ("use_x_prev", C.c_int32),
                ("class_weight", C.c_float), ("kl_weight", C.c_float), ("w_kl_weight", C.c_float),
                ("w_log_var_prior", C.c_float),
                ("gen_noise", C.c_int32), ("do_backward", C.c_int32), ("accumulate", C.c_int32),
                ("gemm_algo", C.c_int32), ("x_shift", C.c_int32), ("gemm_algo_tc_lstm_min", C.c_int32), ("overlap_wgrad", C.c_int32),
                ("y_shift", C.c_int32), ("pair_bwd", C.c_int32),
                ("seed", C.c_uint64)]


class clv_gemm_args(C.Structure):
    _fields_ = [("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
                ("A", C.c_void_p), ("lda", C.c_int64), ("a_u8", C.c_int32), ("a_kmajor", C.c_int32),
                ("a_off", C.c_void_p), ("a_grp", C.c_int32), ("a_shift", C.c_int32),
                ("a_row_delta", C.c_int32), ("a_skip_grp", C.c_int32),
                ("Bm", C.c_void_p), ("ldb", C.c_int64), ("b_nmajor", C.c_int32),
                ("C", C.c_void_p), ("ldc", C.c_int64),
                ("bias", C.c_void_p),
                ("rowadd", C.c_void_p), ("ldra", C.c_int64), ("ra_grp", C.c_int32),
                ("relu_mask", C.c_void_p), ("ldmask", C.c_int64),
                ("relu", C.c_int32), ("accumulate", C.c_int32), ("split_k", C.c_int32)]


_P, _I32, _I64, _U64, _F, _D = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_double
_CFG = C.POINTER(clv_cfg)

# name -> (restype, argtypes); every symbol declared in include/clv_b200.h
PROTOTYPES = {
    "clv_version": (C.c_int, []),
    "clv_launch_count": (_I64, []),
    "clv_runtime_init": (C.c_int, []),
    "clv_error_string": (C.c_char_p, [C.c_int]),
    "clv_param_layout": (_I64, [_CFG, C.POINTER(_I64), C.POINTER(_I32), C.POINTER(_I32)]),
    "clv_gemm": (C.c_int, [C.POINTER(clv_gemm_args), _P]),
    "clv_inproj_tc_scratch_bytes": (_I64, []),
    "clv_inproj_tc": (C.c_int, [_P, _P, _I32, _I32, _I32, _P, _I64, _I32, _P, _P, _I64, _I64, _P, _I64, _I32, _P]),
    "clv_lstm_wgrad_tc": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _P, _P, _I32, _P, _P, _P, _I64, _I32, _P]),
    "clv_bias_act": (C.c_int, [_P, _I64, _I32, _I32, _P, _I32, _P]),
    "clv_colsum": (C.c_int, [_P, _I64, _I32, _I32, _P, _I32, _P]),
    "clv_logitnormal_fwd": (C.c_int, [_P, _I64, _P, _P, _P, _P, _I32, _I32, _F, _F, _I32, _U64, _P, _P]),
    "clv_logitnormal_bwd": (C.c_int, [_P, _I64, _P, _P, _P, _P, _P, _I32, _I32, _F, _F, _F, _P]),
    "clv_gauss_heads_fwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _F, _I32,
                                      _U64, _P, _P]),
    "clv_gauss_heads_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _F,
                                      _I32, _P]),
    "clv_lstm_fwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _P]),
    "clv_lstm_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _P]),
    "clv_lstm_fwd_fused": (C.c_int, [_P, _I32, _P, _P, _P, _P, _I32, _P, _P, _I32, _P, _P, _I32, _I32, _I32, _P]),
    "clv_lstm_bwd_fused": (C.c_int, [_P, _P, _P, _P, _P, _P, _I32, _P, _I32, _P, _I32, _P, _I32, _I32, _I32, _P]),
    "clv_lstm_bwd_heads": (C.c_int, [_P, _P, _P, _P, _P, _P, _I32, _P, _I32, _P, _I32, _P, _P, _P, _F, _P, _P, _P,
                                     _P, _I32, _I32, _I32, _I32, _P]),
    "clv_lstm_pair_fwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I32, _P, _P, _P, _P, _P, _P, _P, _I32, _P, _P, _P, _P,
                                    _P, _P, _P, _P, _F, _I32, _U64, _P, _I32, _I32, _I32, _I32, _P]),
    "clv_vae_fused_step": (C.c_int, [_CFG, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "clv_lstm_pair_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _F, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                    _I32, _I32, _I32, _I32, _I32, _P]),
    "clv_lstm_fwd_tc_scratch_bytes": (_I64, []),
    "clv_lstm_fwd_tc": (C.c_int, [_P, _P, _P, _P, _I32, _P, _P, _P, _I32, _I32, _I32, _P]),
    "clv_xhead_fwd_bwd": (C.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _P, _P, _P, _I64, _I32, _I32, _F, _I32, _P]),
    "clv_xhead_tc_scratch_bytes": (C.c_int64, []),
    "clv_xhead_tc": (C.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _F, _P]),
    "clv_keyenc_fwd": (C.c_int, [_P, _P, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I32, _I32,
                                 _F, _F, _I32, _U64, _P, _P]),
    "clv_keyenc_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _F, _F, _F, _P]),
    "clv_keyenc_bwd_full": (C.c_int, [_P, _P, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                      _I32, _I32, _F, _F, _F, _P]),
    "clv_bernoulli_ce_fwd_bwd": (C.c_int, [_P, _P, _P, _I32, _I32, _P, _I64, _I32, _F, _I32, _P]),
    "clv_adamwn_state_floats": (_I64, [_CFG]),
    "clv_adamwn_init": (C.c_int, [_CFG, _P, _P]),
    "clv_adamwn_step": (C.c_int, [_CFG, _P, _P, _P, _D, _D, _D, _D, _D, _I32, _P]),
    "clv_adamwn_step_range": (C.c_int, [_CFG, _P, _P, _P, _D, _D, _D, _D, _D, _I32, _I32, _I32, _I32, _P, _P]),
    "clv_adamwn_step_p2p": (C.c_int, [_CFG, _P, _P, _I32, _P, _P, _P, _D, _D, _D, _D, _I32, _P]),
    "clv_step_begin": (C.c_int, [_P, _P, _I32, _I32, _P]),
    "clv_workspace_bytes": (_I64, [_CFG]),
    "clv_workspace_offset": (_I64, [_CFG, C.c_char_p]),
    "clv_train_step": (C.c_int, [_CFG, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _P]),
    "clv_train_step_opt": (C.c_int, [_CFG, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _P, _P]),
    "clv_vrnn_sample": (C.c_int, [_CFG, _P, _P, _P, _P, _P, _I32, _I32, _P, _P, _P, _U64, _I64, _I32,
                                  _P, _P, _P]),
    "clv_vrnn_sample_bits": (C.c_int, [_CFG, _P, _P, _P, _P, _P, _I32, _I32, _P, _P, _P, _U64, _I64, _I32,
                                       _P, _P, _P]),
    "clv_vae_sample": (C.c_int, [_CFG, _P, _P, _I32, _P, _P, _P, _U64, _I64, _I32, _I32, _P, _P, _P]),
    "clv_chunk_mean": (C.c_int, [_P, _P, _I32, _I32, _I32, _P]),
    "clv_p2p_flag_ints": (C.c_int, []),
    "clv_p2p_signal": (C.c_int, [_P, _P, _CFG, _I32, _P]),
    "clv_p2p_wait_done": (C.c_int, [_P, _P, _CFG, _P]),
    "clv_p2p_allreduce": (C.c_int, [_P, _P, _CFG, _I64, _I64, _I32, _I32, _P]),
    "clv_p2p_allreduce_blocks": (C.c_int, [_I64]),
    "clv_adamwn_range_blocks": (C.c_int, [_CFG, _I32, _I32, _I32]),
    "clv_adamwn_step_range_p2p": (C.c_int, [_CFG, _P, _P, _P, _D, _D, _D, _D, _I32, _I32, _I32, _I32, _I32, _P, _P]),
    "clv_fp32_peak_probe": (C.c_int, [_I32, _P, _I64, C.POINTER(C.c_double), _P]),
}

class clv_adam_args(C.Structure):
    """include/clv_b200.h: clv_adam_args (optimizer settings of clv_train_step_opt)."""
    _fields_ = [("state", C.c_void_p), ("lr", C.c_double), ("beta_1", C.c_double), ("beta_2", C.c_double),
                ("epsilon", C.c_double), ("grad_scale", C.c_double), ("weightnorm", C.c_int32),
                ("loss_mirror", C.c_void_p), ("exchange", C.c_void_p), ("exchange_user", C.c_void_p),
                ("p2p", C.c_void_p)]


class clv_p2p_args(C.Structure):
    """include/clv_b200.h: clv_p2p_args (peer-memory data parallelism, hand-shake inside the kernels)."""
    _fields_ = [("peer_grads", C.c_void_p), ("peer_flags", C.c_void_p), ("n_peers", C.c_int32), ("rank", C.c_int32),
                ("gsum", C.c_void_p), ("loss_out", C.c_void_p), ("form", C.c_int32),
                ("mc_grads", C.c_void_p), ("mc_gsum", C.c_void_p)]


# include/clv_b200.h: clv_exchange_fn(user, buf, count, stream) -> int
EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)


_lib = None


def lib():
    """Load libclv_b200.so (once).  Raises ClvError if it was not built -- run
    `python -c "import __graft_entry__ as g; g.build()"` or classifying-vae-lstm_b200/csrc/build.sh."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ClvError("libclv_b200.so not built at %s; there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc, what=""):
    if rc < 0:
        raise ClvError("%s failed: %s (code %d)" % (what or "clv call", lib().clv_error_string(int(rc)).decode(), rc))
    return rc


def ptr(t):
    """Raw device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def make_cfg(model, B, L, D, H, Z, C_, use_x_prev, Hc=0, B_global=None, class_weight=1.0,
             kl_weight=1.0, w_kl_weight=1.0, w_log_var_prior=0.0, gen_noise=0, do_backward=1,
             accumulate=0, gemm_algo=0, seed=0, x_shift=0, overlap_wgrad=0, tc_lstm_min=0, y_shift=0, pair_bwd=0):
    return clv_cfg(model=model, B=B, B_global=B if B_global is None else B_global, L=L, D=D, H=H,
                   Hc=Hc, Z=Z, C=C_, use_x_prev=int(bool(use_x_prev)), class_weight=class_weight,
                   kl_weight=kl_weight, w_kl_weight=w_kl_weight, w_log_var_prior=w_log_var_prior,
                   gen_noise=gen_noise, do_backward=do_backward, accumulate=accumulate,
                   gemm_algo=gemm_algo, x_shift=x_shift, gemm_algo_tc_lstm_min=tc_lstm_min,
                   overlap_wgrad=overlap_wgrad, y_shift=y_shift, pair_bwd=pair_bwd, seed=seed)


def param_layout(cfg):
    offs = (_I64 * CLV_N_TENSORS)()
    rows = (_I32 * CLV_N_TENSORS)()
    cols = (_I32 * CLV_N_TENSORS)()
    P = check(lib().clv_param_layout(C.byref(cfg), offs, rows, cols), "clv_param_layout")
    return int(P), list(offs), list(rows), list(cols)

"""
Small device-side building blocks for the sampler sub-models' `.predict()` (the Keras objects that
make_w_encoder / make_z_encoder / make_decoder return in the reference, cl_vrnn/model.py:98-162,
cl_vae/model.py:76-128).  Every op is a libclv_b200 call (clv_gemm, clv_bias_act, clv_lstm_fwd);
PyTorch only allocates.  These are the one-call-per-step API the reference's Python loop uses -- the
throughput path is the persistent sampler kernel (generate_samples).
"""
import ctypes as C
import numpy as np
import torch

from . import _lib
from ._lib import lib, check, ptr
from .engine import _stream


def as_dev_f32(a, dev):
    t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32)))
    return t.to(device=dev, dtype=torch.float32).contiguous()


def dense(x, kernel, bias=None, act=0, out=None, accumulate=False):
    """out[M,N] = act(x[M,K] @ kernel[K,N] + bias); act: 0 none, 1 relu, 2 sigmoid.  `kernel` may be a
    row-slice view of a parameter tensor (contiguous rows)."""
    M, K = x.shape
    N = kernel.shape[1]
    assert kernel.shape[0] == K and kernel.stride(1) == 1 and x.stride(1) == 1
    if out is None:
        out = torch.empty(M, N, device=x.device, dtype=torch.float32)
    fuse = act in (0, 1) and not accumulate
    a = _lib.clv_gemm_args(M=M, N=N, K=K, A=x.data_ptr(), lda=x.stride(0), a_kmajor=1,
                           Bm=kernel.data_ptr(), ldb=kernel.stride(0), b_nmajor=1, C=out.data_ptr(),
                           ldc=out.stride(0), bias=(bias.data_ptr() if (bias is not None and fuse) else None),
                           relu=int(act == 1 and fuse), accumulate=int(accumulate), split_k=1)
    check(lib().clv_gemm(C.byref(a), _stream()), "clv_gemm")
    if not fuse:
        check(lib().clv_bias_act(ptr(out), out.stride(0), M, N, ptr(bias) if bias is not None else None,
                                 int(act), _stream()), "clv_bias_act")
    return out


class StatefulLSTM:
    """A Keras `LSTM(stateful=True, return_sequences=True)` layer driven one call at a time:
    input projection through clv_gemm, recurrence through clv_lstm_fwd with (h0, c0) carried over."""
    def __init__(self, kernel, rkernel, bias):
        self.kernel, self.rkernel, self.bias = kernel, rkernel, bias
        self.H = rkernel.shape[0]
        self.h = self.c = None

    def reset_states(self):
        self.h = self.c = None

    def __call__(self, xin):
        """xin [S, L, In] float32 device -> all h_t [S, L, H]"""
        S, L, In = xin.shape
        G = 4 * self.H
        gates = dense(xin.reshape(S * L, In), self.kernel, self.bias).reshape(S, L, G)
        hs = torch.empty(S, L, self.H, device=xin.device)
        cs = torch.empty(S, L, self.H, device=xin.device)
        if self.h is not None and self.h.shape[0] != S:
            raise ValueError("stateful LSTM: batch size changed from %d to %d without reset_states()" % (self.h.shape[0], S))
        check(lib().clv_lstm_fwd(ptr(gates), ptr(self.rkernel), ptr(hs), ptr(cs), ptr(self.h), ptr(self.c),
                                 S, L, self.H, _stream()), "clv_lstm_fwd")
        self.h, self.c = hs[:, -1].contiguous(), cs[:, -1].contiguous()
        return hs

"""Shared helpers for the parity tests (the oracle is the checker, never the thing shipped)."""
import numpy as np
import torch

import clvae_b200  # noqa: F401  (alias of classifying-vae-lstm_b200)
from oracle import clv_oracle as O


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def make_vrnn_case(seed, B, L, C=10, Z=2, use_x_prev=True, D=88, H=88, density=0.08):
    rng = np.random.default_rng(seed)
    p = O.init_vrnn_params(rng, L, D, H, Z, C, use_x_prev)
    # make biases / heads non-trivial so every gradient path is exercised
    for k in p:
        if k.endswith(".bias"):
            p[k] = p[k] + torch.tensor(rng.normal(0, 0.1, p[k].shape))
    W = L + 1 if use_x_prev else L
    win = O.synth_rolls(rng, B, W, D, density)
    labels = rng.integers(0, C, B).astype(np.int32)
    eps_w = rng.standard_normal((B, C - 1))
    eps_z = rng.standard_normal((B, L, Z))
    return dict(p=p, win=win, labels=labels, eps_w=eps_w, eps_z=eps_z, B=B, L=L, C=C, Z=Z, D=D, H=H,
                use_x_prev=use_x_prev)


def oracle_vrnn(case, dtype=torch.float64, **kw):
    p = {k: v.to(dtype) for k, v in case["p"].items()}
    win = torch.tensor(case["win"], dtype=dtype)
    if case["use_x_prev"]:
        X, Xp = win[:, 1:], win[:, :-1]
    else:
        X, Xp = win, None
    return O.vrnn_loss_and_grads(p, X, Xp, O.one_hot(case["labels"], case["C"], dtype),
                                 torch.tensor(case["eps_w"], dtype=dtype),
                                 torch.tensor(case["eps_z"], dtype=dtype), case["C"],
                                 case["use_x_prev"], **kw)


def make_vae_case(seed, B, C=2, Z=4, use_x_prev=True, D=88, H=88, Hc=88, density=0.08):
    rng = np.random.default_rng(seed)
    p = O.init_vae_params(rng, D, H, Z, Hc, C, use_x_prev)
    for k in p:
        if k.endswith(".bias"):
            p[k] = p[k] + torch.tensor(rng.normal(0, 0.1, p[k].shape))
    W = 2 if use_x_prev else 1
    win = O.synth_rolls(rng, B, W, D, density)
    labels = rng.integers(0, C, B).astype(np.int32)
    return dict(p=p, win=win, labels=labels, eps_w=rng.standard_normal((B, C - 1)),
                eps_z=rng.standard_normal((B, Z)), B=B, C=C, Z=Z, D=D, H=H, Hc=Hc,
                use_x_prev=use_x_prev)


def oracle_vae(case, dtype=torch.float64, **kw):
    p = {k: v.to(dtype) for k, v in case["p"].items()}
    win = torch.tensor(case["win"], dtype=dtype)
    if case["use_x_prev"]:
        x, xp = win[:, 1], win[:, 0]
    else:
        x, xp = win[:, 0], None
    return O.vae_loss_and_grads(p, x, xp, O.one_hot(case["labels"], case["C"], dtype),
                                torch.tensor(case["eps_w"], dtype=dtype),
                                torch.tensor(case["eps_z"], dtype=dtype), case["C"],
                                case["use_x_prev"], **kw)


def engine_for(case, model, **kw):
    from clvae_b200.engine import Engine
    e = Engine(model, case["B"], L=case.get("L", 1), D=case["D"], H=case["H"], Z=case["Z"],
               n_classes=case["C"], use_x_prev=case["use_x_prev"], Hc=case.get("Hc", 88), **kw)
    e.set_params({k: v.numpy() for k, v in case["p"].items()})
    e.stage_windows(torch.tensor(case["win"]).cuda(), torch.tensor(case["labels"]).cuda())
    e.eps_w.copy_(torch.tensor(case["eps_w"], dtype=torch.float32).reshape(-1))
    e.eps_z.copy_(torch.tensor(case["eps_z"], dtype=torch.float32).reshape(-1))
    return e


# ---------------------------------------------------------------------------------- golden fixtures
def golden_weight(seed, name, shape, scale=1.0):
    """Weights of the tests/golden/*_train.npz / *_sampler.npz cases: regenerated from (seed, tensor
    name) instead of being stored (random float32 does not compress); the fixture keeps a SHA-256 of
    each tensor so a drift of the numpy stream is caught.  Non-default values so that every branch is
    exercised (hard-sigmoid clips, relu on/off, non-zero biases); float32-representable."""
    import zlib
    rng = np.random.default_rng([int(seed), zlib.crc32(name.encode())])
    layer, wname = name.rsplit(".", 1)
    if wname == "bias":
        s = 0.25
    else:
        s = scale * 0.9 / np.sqrt(max(shape[0], 1))
        if layer in ("Z_mean", "Z_log_var", "z_mean", "z_log_var", "Wargs", "w_mean", "w_log_var"):
            s *= 0.5
    return rng.normal(0, s, shape).astype(np.float32)


def sha256(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_golden(fname):
    import json, os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", fname))
    cases = {}
    for k in z.files:
        case, rest = k.split("/", 1)
        cases.setdefault(case, {})[rest] = z[k]
    for c in cases.values():
        c["cfg"] = json.loads(str(c["cfg"]))
    return cases


def golden_params(case, names_shapes, prefix=""):
    """regenerate + verify the weights of one golden case; names_shapes: [(tensor name, shape)]"""
    cfg = case["cfg"]
    out = {}
    for name, shape in names_shapes:
        w = golden_weight(cfg["wseed"], prefix + name, shape, cfg["wscale"])
        assert sha256(w) == cfg["wsha"][prefix + name], "numpy stream drift: regenerate tests/golden"
        out[name] = w
    return out

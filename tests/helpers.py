"""Shared test helpers: golden-fixture access and tolerance checks."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_cache = {}


def golden(name):
    if name not in _cache:
        _cache[name] = np.load(os.path.join(GOLDEN, name + ".npz"))
    return _cache[name]


def case(npz, prefix):
    """All arrays below `prefix/` as {relative key: torch tensor}; nested param/ and grad_param/ dicts."""
    out, params, gparams = {}, {}, {}
    pre = prefix + "/"
    for k in npz.files:
        if not k.startswith(pre):
            continue
        rel = k[len(pre):]
        arr = npz[k]
        val = torch.from_numpy(arr) if arr.dtype.kind in "fiub" and arr.ndim > 0 else arr
        if rel.startswith("param/"):
            params[rel[6:]] = val
        elif rel.startswith("grad_param/"):
            gparams[rel[11:]] = val
        elif "/" not in rel:
            out[rel] = val
    out["param"], out["grad_param"] = params, gparams
    return out


def json_blob(npz, key="kwargs_json"):
    return json.loads(bytes(npz[key]).decode())


def none_str(x):
    x = str(x)
    return None if x == "None" else x


def rel_err(a, b):
    """max |a-b| / max(|b|, tiny): the scale-relative error used for all fp tolerances."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    if b.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def assert_close(a, b, tol, what="", atol=0.0):
    """|a-b|_max <= tol * |b|_max + atol  (atol only for quantities that are analytically zero)."""
    assert tuple(a.shape) == tuple(b.shape), f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    a64, b64 = a.detach().double().cpu(), b.detach().double().cpu()
    if b64.numel() == 0:
        return
    assert torch.isfinite(a64).all(), f"{what}: non-finite values"
    err, scale = float((a64 - b64).abs().max()), float(b64.abs().max())
    assert err <= tol * scale + atol, f"{what}: abs err {err:.3e}, scale {scale:.3e}, rel {err / max(scale, 1e-30):.3e} > {tol:.1e}"

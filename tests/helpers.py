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


def assert_rows_close(a, b, tol, what="", atol=0.0, max_bad_rows=0.01):
    """Per-row tensors of a BACKWARD pass compared across arithmetic orders (fp32 vs fp64, or two kernel families): a ReLU
    pre-activation within round-off of zero can take the other branch in one of them, which switches that row's whole term --
    a discontinuity of the function, not an arithmetic error.  At most `max_bad_rows` of the rows may hold an entry beyond
    `tol` (scale-relative), and no entry may be off by more than the tensor's scale."""
    assert tuple(a.shape) == tuple(b.shape), f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    a64, b64 = a.detach().double().cpu(), b.detach().double().cpu()
    if b64.numel() == 0:
        return
    assert torch.isfinite(a64).all(), f"{what}: non-finite values"
    scale = float(b64.abs().max())
    diff = (a64 - b64).abs()
    bad = (diff.reshape(diff.shape[0], -1) > tol * scale + atol).any(dim=1)
    assert float(bad.double().mean()) <= max_bad_rows, f"{what}: {int(bad.sum())} of {bad.numel()} rows disagree beyond {tol:.1e}"
    assert float(diff.max()) <= scale + atol, f"{what}: max diff {float(diff.max()):.3e} at scale {scale:.3e}"


def assert_param_grads_close(got, want, what="", each=1e-2, together=1e-3, atol=1e-5):
    """Parameter gradients ({name: tensor}) are signed sums over all rows with heavy cancellation, so one flipped ReLU term (see
    `assert_rows_close`) shows at up to ~1e-3 of a tensor's scale: each tensor within `each`, all of them together within
    `together` (L2-relative)."""
    fa, fb = [], []
    for k, w in want.items():
        g = got[k]
        assert tuple(g.shape) == tuple(w.shape), f"{what} {k}: shape"
        g64, w64 = g.detach().double().cpu(), w.detach().double().cpu()
        assert torch.isfinite(g64).all(), f"{what} {k}: non-finite values"
        err, scale = float((g64 - w64).abs().max()), float(w64.abs().max())
        assert err <= each * scale + atol, f"{what} grad {k}: abs err {err:.3e} at scale {scale:.3e}"
        fa.append(g64.flatten())
        fb.append(w64.flatten())
    fa, fb = torch.cat(fa), torch.cat(fb)
    rel = float((fa - fb).norm() / fb.norm().clamp(min=1e-30))
    assert rel <= together, f"{what}: parameter gradients together, L2-relative {rel:.3e} > {together:.1e}"

"""CPU check of the register-resident GVP / LayerNorm templates (caster_dta_b200/csrc/cgvp_reg.cuh).

The templates are `__host__ __device__`; tests/csrc/reg_harness.cu instantiates them for the host, this file packs
weights with an independent numpy restatement of the packed layout (cgvp_common.cuh) and compares forward values,
input gradients and weight gradients with torch autograd through the oracle (fp64).  No GPU involved; the same
templates are what conv_reg.cu / rows_reg.cu run on the device.
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from oracle import gvp_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ACT = {0: None, 1: "relu", 2: "sigmoid"}
# (si, vi, so, vo, h, sact, vact, gate) -- must mirror the switch in reg_harness.cu
CASES = {
    0: (64, 9, 16, 4, 9, 1, 0, 1), 1: (16, 4, 16, 4, 4, 1, 0, 1), 2: (16, 4, 16, 4, 4, 0, 0, 1),
    3: (37, 3, 16, 4, 4, 0, 0, 1), 4: (33, 1, 32, 1, 1, 0, 0, 1), 5: (16, 4, 64, 8, 8, 1, 0, 1),
    6: (64, 8, 16, 4, 8, 0, 0, 1), 7: (16, 4, 64, 0, 4, 1, 0, 1), 8: (10, 3, 7, 5, 5, 1, 2, 0),
    9: (6, 0, 5, 0, 0, 1, 0, 0), 10: (12, 5, 9, 3, 6, 2, 0, 0),
}


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    out = str(tmp_path_factory.mktemp("h") / "reg_harness.so")
    src = os.path.join(ROOT, "tests", "csrc", "reg_harness.cu")
    subprocess.run(["nvcc", "-O1", "-std=c++17", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets", "-shared",
                    "-Xcompiler", "-fPIC", "-o", out, src], check=True)
    return C.CDLL(out)


def pad4(x):
    return (x + 3) // 4 * 4


def pack(si, vi, so, vo, h, gate, p):
    """numpy restatement of pack_kernel (pack.cu) / make_gvp_p (cgvp_common.cuh).  Returns (packed, fwd_floats)."""
    h = h if vi else 0
    gate = bool(gate and vi and vo)
    vip, hp, sop, vop = pad4(vi), pad4(h), pad4(so), pad4(vo)
    ksd = si + h
    ksp, ksvp, ksdp = pad4(ksd + 1), pad4(so + 1), pad4(ksd)
    wh_t = np.zeros((vip, hp), np.float32)
    ws_t = np.zeros((ksp, sop), np.float32)
    wv_t = np.zeros((hp if vi else 0, vop), np.float32)
    wsv_t = np.zeros((ksvp if gate else 0, vop), np.float32)
    wh_b = np.zeros((hp, vip), np.float32)
    ws_b = np.zeros((sop, ksdp), np.float32)
    wv_b = np.zeros((vop, hp if vi else 0), np.float32)
    wsv_b = np.zeros((vop if gate else 0, sop), np.float32)
    ws, bs = p["ws.weight"].numpy(), p["ws.bias"].numpy()
    ws_t[:ksd, :so] = ws.T
    ws_t[ksd, :so] = bs
    ws_b[:so, :ksd] = ws
    if vi:
        wh = p["wh.weight"].numpy()
        wh_t[:vi, :h] = wh.T
        wh_b[:h, :vi] = wh
        if vo:
            wv = p["wv.weight"].numpy()
            wv_t[:h, :vo] = wv.T
            wv_b[:vo, :h] = wv
            if gate:
                wsv, bg = p["wsv.weight"].numpy(), p["wsv.bias"].numpy()
                wsv_t[:so, :vo] = wsv.T
                wsv_t[so, :vo] = bg
                wsv_b[:vo, :so] = wsv
    fwd = [wh_t, ws_t, wv_t, wsv_t]
    bwd = [wh_b, ws_b, wv_b, wsv_b]
    flat = np.concatenate([a.ravel() for a in fwd + bwd]).astype(np.float32)
    return flat, sum(a.size for a in fwd), [a.shape for a in fwd]


def fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


@pytest.mark.parametrize("which", sorted(CASES) + [100])
def test_gvp_templates_match_oracle(harness, which):
    """which = 100: message GVP 0 through the round-2 split path -- per-node projections of the node scalars
    (gvp_fwd<G, 2, NS, ES>), slice-mode backward (gvp_bwd_ds) finished per node as conv_node_post_kernel does, and the
    stash path (gvp_fwd<G, 1>) -- must give the same values and gradients as the plain GVP."""
    si, vi, so, vo, h, sact, vact, gate = CASES[0 if which == 100 else which]
    gen = torch.Generator().manual_seed(100 + which)
    n = 13
    p = O.init_gvp_params({}, "", (si, vi), (so, vo), h_dim=h or None, vector_gate=bool(gate), gen=gen)
    p = {k: v for k, v in p.items()}
    s = torch.randn(n, si, generator=gen)
    v = torch.randn(n, max(vi, 1), 3, generator=gen)
    if vi:
        v[0] = 0.0                                   # exercises the clamp in _norm_no_nan (zero gradient branch)
    gs = torch.randn(n, so, generator=gen)
    gv = torch.randn(n, max(vo, 1), 3, generator=gen)
    W, fwd_floats, shapes = pack(si, vi, so, vo, h, gate, p)
    W = np.ascontiguousarray(W)
    assert W.ctypes.data % 16 == 0 or True
    Wa = np.zeros(W.size + 4, np.float32)
    off = (-Wa.ctypes.data // 4) % 4                 # 16-byte aligned view
    Wv = Wa[off:off + W.size]
    Wv[:] = W
    out_s = np.zeros((n, so), np.float32)
    out_v = np.zeros((n, max(vo, 1), 3), np.float32)
    d_s = np.zeros((n, si), np.float32)
    d_v = np.zeros((n, max(vi, 1), 3), np.float32)
    G = np.zeros(fwd_floats, np.float32)
    sv, vv = s.numpy().copy(), v[:, :vi].numpy().copy() if vi else np.zeros((n, 0, 3), np.float32)
    gsv = gs.numpy().copy()
    gvv = gv[:, :vo].numpy().copy() if vo else np.zeros((n, 0, 3), np.float32)
    ov = np.zeros((n, vo, 3), np.float32) if vo else np.zeros((1,), np.float32)
    dv = np.zeros((n, vi, 3), np.float32) if vi else np.zeros((1,), np.float32)
    rc = harness.harness_gvp(which, n, fptr(Wv), fptr(sv), fptr(np.ascontiguousarray(vv)) if vi else None, fptr(gsv),
                             fptr(np.ascontiguousarray(gvv)) if vo else None, fptr(out_s), fptr(ov), fptr(d_s), fptr(dv), fptr(G))
    assert rc == 0, "harness failed (or, for the split case, the stash path changed a bit of the output)"
    # oracle in fp64 with autograd
    pd = {k: t.double().requires_grad_(t.numel() > 0) for k, t in p.items()}
    sd = s.double().requires_grad_(True)
    x = sd
    if vi:
        vd = v[:, :vi].double().requires_grad_(True)
        x = (sd, vd)
    y = O.gvp(pd, "", x, scalar_act=ACT[sact], vector_act=ACT[vact], vector_gate=bool(gate))
    if vo:
        ys, yv = y
        loss = (ys * gs.double()).sum() + (yv * gv[:, :vo].double()).sum()
    else:
        ys = y
        loss = (ys * gs.double()).sum()
    loss.backward()
    tol = 2e-5

    def close(a, b, what):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        scale = max(np.abs(b).max(), 1e-6)
        assert np.abs(a - b).max() / scale < tol, f"{what}: {np.abs(a - b).max() / scale}"

    close(out_s, ys.detach().numpy(), "s")
    if vo:
        close(ov, yv.detach().numpy(), "V")
    close(d_s, sd.grad.numpy(), "ds")
    if vi:
        close(dv, vd.grad.numpy(), "dV")
    # weight gradients, read back from the packed (forward) layout
    h_ = h if vi else 0
    ksd = si + h_
    o = 0
    wh_t = G[o:o + shapes[0][0] * shapes[0][1]].reshape(shapes[0]); o += wh_t.size
    ws_t = G[o:o + shapes[1][0] * shapes[1][1]].reshape(shapes[1]); o += ws_t.size
    wv_t = G[o:o + shapes[2][0] * shapes[2][1]].reshape(shapes[2]); o += wv_t.size
    wsv_t = G[o:o + shapes[3][0] * shapes[3][1]].reshape(shapes[3])
    close(ws_t[:ksd, :so].T, pd["ws.weight"].grad.numpy(), "dws")
    close(ws_t[ksd, :so], pd["ws.bias"].grad.numpy(), "dbs")
    if vi:
        close(wh_t[:vi, :h_].T, pd["wh.weight"].grad.numpy(), "dwh")
        if vo:
            close(wv_t[:h_, :vo].T, pd["wv.weight"].grad.numpy(), "dwv")
            if gate:
                close(wsv_t[:so, :vo].T, pd["wsv.weight"].grad.numpy(), "dwsv")
                close(wsv_t[so, :vo], pd["wsv.bias"].grad.numpy(), "dbg")


@pytest.mark.parametrize("which,S,Cv", [(0, 16, 4), (1, 32, 1), (2, 7, 0)])
def test_layer_norm_templates_match_oracle(harness, which, S, Cv):
    gen = torch.Generator().manual_seed(7 + which)
    n = 11
    p = O.init_layer_norm_params({}, "", S, gen=gen)
    s = torch.randn(n, S, generator=gen)
    v = torch.randn(n, max(Cv, 1), 3, generator=gen)
    if Cv:
        v[1] = 0.0
    gs, gv = torch.randn(n, S, generator=gen), torch.randn(n, max(Cv, 1), 3, generator=gen)
    w, b = p["scalar_norm.weight"].numpy().copy(), p["scalar_norm.bias"].numpy().copy()
    so, ds = np.zeros((n, S), np.float32), np.zeros((n, S), np.float32)
    vo, dv = np.zeros((n, max(Cv, 1), 3), np.float32), np.zeros((n, max(Cv, 1), 3), np.float32)
    dw, db = np.zeros(S, np.float32), np.zeros(S, np.float32)
    vv, gvv = np.ascontiguousarray(v[:, :max(Cv, 1)].numpy()), np.ascontiguousarray(gv.numpy())
    rc = harness.harness_ln(which, n, fptr(w), fptr(b), fptr(s.numpy().copy()), fptr(vv), fptr(gs.numpy().copy()), fptr(gvv),
                            fptr(so), fptr(vo), fptr(ds), fptr(dv), fptr(dw), fptr(db))
    assert rc == 0
    pd = {k: t.double().requires_grad_(True) for k, t in p.items()}
    sd = s.double().requires_grad_(True)
    if Cv:
        vd = v.double().requires_grad_(True)
        ys, yv = O.layer_norm(pd, "", (sd, vd))
        ((ys * gs.double()).sum() + (yv * gv.double()).sum()).backward()
    else:
        ys = O.layer_norm(pd, "", sd)
        (ys * gs.double()).sum().backward()

    def close(a, b_, what, tol=2e-5):
        a, b_ = np.asarray(a, np.float64), np.asarray(b_, np.float64)
        assert np.abs(a - b_).max() / max(np.abs(b_).max(), 1e-6) < tol, what

    close(so, ys.detach().numpy(), "s")
    close(ds, sd.grad.numpy(), "ds")
    close(dw, pd["scalar_norm.weight"].grad.numpy(), "dw")
    close(db, pd["scalar_norm.bias"].grad.numpy(), "db")
    if Cv:
        close(vo, yv.detach().numpy(), "V")
        close(dv, vd.grad.numpy(), "dV")

"""ctypes binding of `libcastergvp.so` (the C ABI declared in `include/castergvp.h`).

The library is built in-tree by `__graft_entry__.build()` / `make -C caster_dta_b200/csrc`.  There is NO fallback:
if the shared object is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os

MAX_CHAIN = 4
ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2
AGGR_SUM, AGGR_MEAN = 0, 1

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcastergvp.so")

c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class GvpDesc(C.Structure):
    _fields_ = [("si", C.c_int32), ("vi", C.c_int32), ("so", C.c_int32), ("vo", C.c_int32), ("h", C.c_int32),
                ("scalar_act", C.c_int32), ("vector_act", C.c_int32), ("vector_gate", C.c_int32)]


class GvpWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("wh", "ws", "bs", "wv", "wsv", "bg")]


class GvpGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("wh", "ws", "bs", "wv", "wsv", "bg")]


class Plan(C.Structure):
    _fields_ = [("num_edges", C.c_int64), ("num_nodes", C.c_int64)] + \
               [(n, C.c_void_p) for n in ("perm", "src", "dst", "rowptr", "sperm", "srowptr")]


class ConvDesc(C.Structure):
    _fields_ = [("ns", C.c_int32), ("nv", C.c_int32), ("es", C.c_int32), ("ev", C.c_int32), ("n_gvp", C.c_int32),
                ("gvp", GvpDesc * MAX_CHAIN), ("aggr", C.c_int32), ("edge_sorted", C.c_int32)]


class RowDesc(C.Structure):
    _fields_ = [("in_s", C.c_int32), ("in_v", C.c_int32), ("onehot", C.c_int32), ("has_residual_in", C.c_int32),
                ("pre_norm", C.c_int32), ("n_gvp", C.c_int32), ("gvp", GvpDesc * MAX_CHAIN),
                ("post_residual", C.c_int32), ("post_norm", C.c_int32)]


class RowArgs(C.Structure):
    _fields_ = [("rows", C.c_int64)] + [(n, C.c_void_p) for n in (
        "in_s", "in_v", "types", "in_index", "h_s", "h_v", "mask0_s", "mask0_v", "mask1_s", "mask1_v",
        "ln0_w", "ln0_b", "ln1_w", "ln1_b", "h_packed", "out_s", "out_v", "stash")]


class RowGradArgs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "d_out_s", "d_out_v", "d_in_s", "d_in_v", "d_h_s", "d_h_v", "d_ln0_w", "d_ln0_b", "d_ln1_w", "d_ln1_b",
        "h_packed_grads")]


# every entry point of include/castergvp.h: name -> (restype, argtypes)
SIGNATURES = {
    "cgvp_last_error": (C.c_char_p, []),
    "cgvp_version": (C.c_int32, []),
    "cgvp_sm_count": (C.c_int32, []),
    "cgvp_set_fast_paths": (C.c_int32, [C.c_int32]),
    "cgvp_set_tensor_cores": (C.c_int32, [C.c_int32]),
    "cgvp_set_wide_gemm": (C.c_int32, [C.c_int32]),
    "cgvp_profile_enable": (C.c_int32, [C.c_int32]),
    "cgvp_profile_collect": (C.c_int32, [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "cgvp_gvp_packed_floats": (C.c_int64, [C.POINTER(GvpDesc)]),
    "cgvp_pack_weights": (C.c_int32, [C.c_int32, C.POINTER(GvpDesc), C.POINTER(GvpWeights), C.POINTER(C.c_void_p),
                                      C.c_void_p]),
    "cgvp_unpack_grads": (C.c_int32, [C.c_int32, C.POINTER(GvpDesc), C.POINTER(C.c_void_p), C.POINTER(GvpGrads),
                                      C.c_void_p]),
    "cgvp_plan_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int64]),
    "cgvp_plan_build": (C.c_int32, [C.c_void_p, C.POINTER(Plan), C.c_void_p, C.c_int64, C.c_void_p]),
    "cgvp_gather_message_input": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p]),
    "cgvp_segment_reduce": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                        C.c_void_p, C.c_void_p]),
    "cgvp_conv_workspace_bytes": (C.c_int64, [C.POINTER(ConvDesc), C.c_int64, C.c_int64, C.c_int32]),
    "cgvp_conv_fwd": (C.c_int32, [C.POINTER(ConvDesc), C.POINTER(Plan), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "cgvp_conv_bwd": (C.c_int32, [C.POINTER(ConvDesc), C.POINTER(Plan), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.c_void_p, C.c_int64, C.c_void_p]),
    "cgvp_conv_stash_bytes": (C.c_int64, [C.POINTER(ConvDesc), C.c_int64]),
    "cgvp_conv_fwd_stash": (C.c_int32, [C.POINTER(ConvDesc), C.POINTER(Plan), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "cgvp_conv_bwd_stash": (C.c_int32, [C.POINTER(ConvDesc), C.POINTER(Plan), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "cgvp_rows_workspace_bytes": (C.c_int64, [C.POINTER(RowDesc), C.c_int64, C.c_int32]),
    "cgvp_rows_stash_floats": (C.c_int64, [C.POINTER(RowDesc)]),
    "cgvp_rows_fwd": (C.c_int32, [C.POINTER(RowDesc), C.POINTER(RowArgs), C.c_void_p, C.c_int64, C.c_void_p]),
    "cgvp_rows_bwd": (C.c_int32, [C.POINTER(RowDesc), C.POINTER(RowArgs), C.POINTER(RowGradArgs), C.c_void_p,
                                  C.c_int64, C.c_void_p]),
    "cgvp_featurize_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int64]),
    "cgvp_featurize_count": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_double,
                                         C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "cgvp_featurize_fill": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_int32,
                                        C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_int64, C.c_void_p]),
    "cgvp_attn_supported": (C.c_int32, [C.c_int32, C.c_int32]),
    "cgvp_attn_fwd": (C.c_int32, [C.c_void_p] * 6 + [C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_float,
                                  C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cgvp_attn_bwd": (C.c_int32, [C.c_void_p] * 10 + [C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_float,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cgvp_linear_wgrad_supported": (C.c_int32, [C.c_int64, C.c_int32, C.c_int32]),
    "cgvp_linear_wgrad_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int32, C.c_int32]),
    "cgvp_linear_wgrad": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_int64, C.c_void_p]),
    "cgvp_linear_gemm_supported": (C.c_int32, [C.c_int64, C.c_int32, C.c_int32]),
    "cgvp_linear_fwd": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "cgvp_linear_dgrad": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "cgvp_layernorm_supported": (C.c_int32, [C.c_int32]),
    "cgvp_layernorm_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int32]),
    "cgvp_layernorm_fwd": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "cgvp_layernorm_bwd": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "cgvp_node_features": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32,
                                       C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
}

KERNEL_IDS = {"conv_fwd": 0, "conv_bwd": 1, "rows_fwd": 2, "rows_bwd": 3, "segment_reduce": 4, "gather": 5,
              "featurize": 6}


def set_fast_paths(on):
    """1 (default): specialised register-resident kernels where compiled in; 0: generic tile kernels only."""
    lib().cgvp_set_fast_paths(int(bool(on)))


TENSOR_CORES = False   # mirror of the library's precision mode for the host-side paths (caster_dta_b200/wide.py)


def set_tensor_cores(on):
    """0 (default): fp32 everywhere; 1: tcgen05 bf16 message GEMMs where compiled in, TF32 library GEMMs in the
    wide-dims GEMM formulation (<= 1e-2 of the reference)."""
    global TENSOR_CORES
    lib().cgvp_set_tensor_cores(int(bool(on)))
    TENSOR_CORES = bool(on)


def profile_enable(on):
    lib().cgvp_profile_enable(int(on))


def profile_collect():
    """{kernel name: (total ms, launches)} for everything recorded since profile_enable(True)."""
    out = {}
    for name, kid in KERNEL_IDS.items():
        ms, n = C.c_double(0), C.c_int64(0)
        lib().cgvp_profile_collect(kid, C.byref(ms), C.byref(n))
        out[name] = (ms.value, n.value)
    return out


_lib = None
LAUNCHES = 0   # number of library calls that enqueue kernels (bench.py reports it)


def lib():
    """Load the shared object once.  Raises if it has not been built -- there is no CPU path."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(or `make -C caster_dta_b200/csrc`). There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


# optional per-entry-point device timing (bench.py): name -> list of (start_event, end_event)
TIMED = None          # set to a set of entry-point names to time, e.g. {"cgvp_conv_bwd"}
EVENTS = {}


def timed_call(what, fn, *args):
    """Call a C-ABI entry point, optionally bracketed by CUDA events on the current stream."""
    if TIMED is not None and what in TIMED:
        import torch
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = fn(*args)
        b.record()
        EVENTS.setdefault(what, []).append((a, b))
    else:
        rc = fn(*args)
    check(rc, what)


def check(rc, what):
    global LAUNCHES
    LAUNCHES += 1
    if rc != 0:
        raise RuntimeError(f"{what} failed (rc={rc}): {lib().cgvp_last_error().decode()}")

"""Residue-graph featurizer on the GPU: drop-in for `compute_residue_edge_features` + `construct_graph`
(`utils/create_protein_features.py:201-357`, `utils/create_graphs.py:6-62`) applied to a batch of proteins.

`residue_graph_batch` returns the tensors a PyG `Batch` of the reference's `Data` objects would hold:
`edge_index [2,E]` int64 with batch-global node ids sorted by (src, dst), `edge_attr = (s [E,32], V [E,1,3])`
and `edge_type [E]` (all zero: the shipped dataset has a single edge type).
"""
import ctypes as C

import torch

from ._lib import lib, check
from .ops import _aligned_ptr, _ptr, _stream, _workspace

THRESH_TYPES = {"dist": 0, "num": 1, "prop": 2}


def residue_graph_batch(res_coords, ptr, edge_thresh=4.0, thresh_type="dist", keep_self_loops=True):
    """res_coords: [N,4,3] (N, CA, C, O) or [N,3] C-alpha coordinates, fp32, CUDA; ptr: [B+1] int64 boundaries."""
    if not res_coords.is_cuda:
        raise RuntimeError("the featurizer needs CUDA tensors (there is no CPU fallback)")
    ca = (res_coords[:, 1, :] if res_coords.dim() == 3 else res_coords).contiguous().float()
    ptr = ptr.to(device=ca.device, dtype=torch.int64).contiguous()
    n, b = int(ca.shape[0]), int(ptr.shape[0]) - 1
    code = THRESH_TYPES[thresh_type]
    max_len = int((ptr[1:] - ptr[:-1]).max()) if b > 0 else 0
    dev = ca.device
    offsets = torch.empty(n + 1, dtype=torch.int64, device=dev)
    ws = _workspace(lib().cgvp_featurize_workspace_bytes(n, max_len), dev)
    wp, wn = _aligned_ptr(ws)
    check(lib().cgvp_featurize_count(_ptr(ca), _ptr(ptr), b, n, max_len, float(edge_thresh), code, int(keep_self_loops),
                                     _ptr(offsets), wp, wn, _stream()), "cgvp_featurize_count")
    e = int(offsets[-1])                      # the one device->host read of the featurizer
    edge_index = torch.empty(2, e, dtype=torch.int64, device=dev)
    edge_s = torch.empty(e, 32, dtype=torch.float32, device=dev)
    edge_v = torch.empty(e, 1, 3, dtype=torch.float32, device=dev)
    check(lib().cgvp_featurize_fill(_ptr(ca), _ptr(ptr), b, n, max_len, float(edge_thresh), code, int(keep_self_loops),
                                    _ptr(offsets), _ptr(edge_index), e, _ptr(edge_s), _ptr(edge_v), wp, wn, _stream()),
          "cgvp_featurize_fill")
    return edge_index, (edge_s, edge_v), torch.zeros(e, dtype=torch.int64, device=dev)

"""Residue-graph featurizer on the GPU: drop-in for `compute_residue_edge_features` + `construct_graph`
(`utils/create_protein_features.py:201-357`, `utils/create_graphs.py:6-62`) applied to a batch of proteins.

`residue_graph_batch` returns the tensors a PyG `Batch` of the reference's `Data` objects would hold:
`edge_index [2,E]` int64 with batch-global node ids sorted by (src, dst), `edge_attr = (s [E,32], V [E,1,3])`
and `edge_type [E]` (all zero: the shipped dataset has a single edge type).
"""
import ctypes as C
import math

import torch

from ._lib import lib, check
from .ops import _aligned_ptr, _ptr, _stream, _workspace

THRESH_TYPES = {"dist": 0, "num": 1, "prop": 2}


def knn_edge_count(lengths, edge_thresh, thresh_type="num", keep_self_loops=True):
    """Edges the 'num' / 'prop' selections produce for proteins of the given lengths (host integers): every residue keeps
    min(k, candidates) neighbours (`utils/create_protein_features.py:304-327`).  Lets a caller size the outputs without
    reading the count back from the device (`residue_graph_batch(..., num_edges=...)`)."""
    if thresh_type not in ("num", "prop"):
        raise ValueError("the edge count of a radius graph depends on the coordinates")
    total = 0
    for n in lengths:
        n = int(n)
        k = int(math.ceil(edge_thresh * n)) if thresh_type == "prop" else int(edge_thresh)
        total += n * max(0, min(k, n - (0 if keep_self_loops else 1)))
    return total


def residue_graph_batch(res_coords, ptr, edge_thresh=4.0, thresh_type="dist", keep_self_loops=True, max_len=None,
                        num_edges=None):
    """res_coords: [N,4,3] (N, CA, C, O) or [N,3] C-alpha coordinates, fp32, CUDA; ptr: [B+1] int64 boundaries.

    `max_len` (an upper bound of the longest protein) and `num_edges` (the exact edge count, see `knn_edge_count`) are
    optional host-side hints: with both given the call makes no device->host read, so it can be captured in a CUDA graph
    whose `res_coords` / `ptr` buffers are refilled before every replay."""
    if not res_coords.is_cuda:
        raise RuntimeError("the featurizer needs CUDA tensors (there is no CPU fallback)")
    ca = (res_coords[:, 1, :] if res_coords.dim() == 3 else res_coords).contiguous().float()
    ptr = ptr.to(device=ca.device, dtype=torch.int64).contiguous()
    n, b = int(ca.shape[0]), int(ptr.shape[0]) - 1
    code = THRESH_TYPES[thresh_type]
    if max_len is None:
        max_len = int((ptr[1:] - ptr[:-1]).max()) if b > 0 else 0
    max_len = int(max_len)
    dev = ca.device
    offsets = torch.empty(n + 1, dtype=torch.int64, device=dev)
    ws = _workspace(lib().cgvp_featurize_workspace_bytes(n, max_len), dev)
    wp, wn = _aligned_ptr(ws)
    check(lib().cgvp_featurize_count(_ptr(ca), _ptr(ptr), b, n, max_len, float(edge_thresh), code, int(keep_self_loops),
                                     _ptr(offsets), wp, wn, _stream()), "cgvp_featurize_count")
    e = int(offsets[-1]) if num_edges is None else int(num_edges)      # the one device->host read of the featurizer
    edge_index = torch.empty(2, e, dtype=torch.int64, device=dev)
    edge_s = torch.empty(e, 32, dtype=torch.float32, device=dev)
    edge_v = torch.empty(e, 1, 3, dtype=torch.float32, device=dev)
    check(lib().cgvp_featurize_fill(_ptr(ca), _ptr(ptr), b, n, max_len, float(edge_thresh), code, int(keep_self_loops),
                                    _ptr(offsets), _ptr(edge_index), e, _ptr(edge_s), _ptr(edge_v), wp, wn, _stream()),
          "cgvp_featurize_fill")
    return edge_index, (edge_s, edge_v), torch.zeros(e, dtype=torch.int64, device=dev)


def residue_node_features(res_coords, ptr, res_idents=None, aa_table=None, add_residue_posenc=False):
    """Drop-in for `compute_residue_node_features(res_coords, res_idents, True, False, add_residue_posenc,
    aa_table is not None)` (`utils/create_protein_features.py:12-198`) on a batch of proteins.

    res_coords: [N,4,3] fp32 CUDA backbone atoms (N, CA, C, O); ptr: [B+1] int64; res_idents: [N] int64 residue type
    ids; aa_table: [num_types, num_props] fp32 property table indexed by those ids (see `aa_property_table`).
    Returns (s [N, 6 + num_props (+16)], V [N,3,3])."""
    if not res_coords.is_cuda:
        raise RuntimeError("the featurizer needs CUDA tensors (there is no CPU fallback)")
    if res_coords.dim() != 3 or res_coords.shape[1] != 4 or res_coords.shape[2] != 3:
        raise ValueError("res_coords must be [N,4,3] (N, CA, C, O)")
    dev = res_coords.device
    x = res_coords.contiguous().float()
    ptr = ptr.to(device=dev, dtype=torch.int64).contiguous()
    n, b = int(x.shape[0]), int(ptr.shape[0]) - 1
    types = props = 0
    ids = tab = None
    if aa_table is not None:
        if res_idents is None:
            raise ValueError("aa_table needs res_idents")
        tab = aa_table.to(device=dev, dtype=torch.float32).contiguous()
        ids = res_idents.to(device=dev, dtype=torch.int64).contiguous()
        types, props = int(tab.shape[0]), int(tab.shape[1])
    out_s = torch.empty(n, 6 + props + (16 if add_residue_posenc else 0), dtype=torch.float32, device=dev)
    out_v = torch.empty(n, 3, 3, dtype=torch.float32, device=dev)
    check(lib().cgvp_node_features(_ptr(x), _ptr(ptr), b, n, _ptr(ids) if ids is not None else None,
                                   _ptr(tab) if tab is not None else None, types, props, int(bool(add_residue_posenc)),
                                   _ptr(out_s), _ptr(out_v), _stream()), "cgvp_node_features")
    return out_s, out_v


AA_PROPERTY_DICTS = ("AA_WEIGHTS", "AA_PKAS", "AA_PKBS", "AA_PKCS", "AA_PKIS", "AA_HYDROPHOB", "AA_ALIPHATIC", "AA_AROMATIC",
                     "AA_ACIDIC", "AA_BASIC", "AA_POLAR_NEUTRAL")


def aa_property_table(pd_maps):
    """[num_types, 11] fp32 table from the reference's own `utils.protein_definitions` module (passed in by the caller):
    row = residue integer id (`PROTEIN_INT_1LETTER_MAP`), columns in the order of `:99-103`."""
    letters = pd_maps.PROTEIN_INT_1LETTER_MAP
    rows = [[getattr(pd_maps, d)[letters[i]] for d in AA_PROPERTY_DICTS] for i in sorted(letters)]
    return torch.tensor(rows, dtype=torch.float32)


def protein_graph_batch(res_coords, ptr, res_idents, aa_table=None, edge_thresh=4.0, thresh_type="dist",
                        keep_self_loops=True, add_residue_posenc=False, max_len=None, num_edges=None):
    """Backbone coordinates -> the keyword arguments of the protein encoder, all on the device:
    `construct_graph` (`utils/create_graphs.py:6-62`) for every protein + `Batch.from_data_list`
    (`dataset/dual_dataset.py:543`).  `model.protein_gnn(**protein_graph_batch(...))` runs the GVP stack."""
    x = residue_node_features(res_coords, ptr, res_idents, aa_table, add_residue_posenc)
    edge_index, eattr, etypes = residue_graph_batch(res_coords, ptr, edge_thresh, thresh_type, keep_self_loops, max_len,
                                                    num_edges)
    dev = res_coords.device
    ptr = ptr.to(device=dev, dtype=torch.int64)
    n = int(res_coords.shape[0])
    batch = torch.searchsorted(ptr[1:].contiguous(), torch.arange(n, device=dev), right=True)
    return dict(x=x, edge_index=edge_index, ntypes=res_idents.to(device=dev, dtype=torch.int64), etypes=etypes,
                eattr=eattr, batch=batch)

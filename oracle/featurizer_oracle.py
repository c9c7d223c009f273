"""CPU oracle (TEST INFRASTRUCTURE) -- numpy restatement of the protein residue-graph featurizer.

Restates `utils/create_protein_features.py:201-385` (edge features + sparsification) followed by
`utils/create_graphs.py:6-62` (dense NaN-masked [n,n,F] -> COO, row-major, fp32) as ONE sparse function:
it never materialises the n x n x 35 array, it enumerates the kept (i, j) pairs and evaluates the features
only there.  Arithmetic order and dtypes follow the reference exactly (see the notes in each step), so the
edge set / edge_index is bit-identical and the features are bit-identical up to libm differences.

Also restates the geometric part of the node featurizer (`utils/create_protein_features.py:27-92`).
"""
import numpy as np

RBF_COUNT = 16
RBF_DMAX = 20.0
POSENC_COUNT = 16


def pairwise_ca_distance(ca32):
    """`create_protein_features.py:225-227`: scipy `pdist` (euclidean) on fp32 coordinates = fp64
    `sqrt(((dx*dx) + dy*dy) + dz*dz)` with every op separately rounded (no FMA), dx = fp64(a) - fp64(b)."""
    c = ca32.astype(np.float64)
    d = c[:, None, :] - c[None, :, :]
    sq = d * d
    return np.sqrt((sq[..., 0] + sq[..., 1]) + sq[..., 2])


def select_edges(dist, edge_thresh, thresh_type, keep_self_loops):
    """Boolean [n,n] keep-mask; row = source i, column = target j (`:279-327`).

    'dist': d <= thresh (diagonal kept iff keep_self_loops, since d_ii = 0).
    'num' / 'prop': for each ROW the k smallest distances (`np.argsort(...)[:, :k]`, :319); the diagonal is
    NaN'd first when self loops are dropped (:279-282) and NaN sorts last.  The reference's argsort is the
    default (unstable) kind; only the SET matters downstream and ties are broken here by column index.
    """
    n = dist.shape[0]
    d = dist.copy()
    if not keep_self_loops:
        d[np.arange(n), np.arange(n)] = np.nan
    if edge_thresh is None:
        keep = ~np.isnan(d)
    elif thresh_type == "dist":
        with np.errstate(invalid="ignore"):
            keep = d <= edge_thresh
    else:
        if thresh_type == "prop":
            k = int(np.ceil(edge_thresh * n))
        elif thresh_type == "num":
            k = int(edge_thresh)
        else:
            raise ValueError(thresh_type)
        idx = np.argsort(d, axis=-1, kind="stable")[:, :k]
        keep = np.zeros((n, n), dtype=bool)
        keep[np.arange(n)[:, None], idx] = True
        # the reference copies features (possibly NaN on the diagonal) for the selected columns, and
        # construct_graph drops all-NaN rows: a selected NaN-diagonal entry is therefore NOT an edge.
        keep &= ~np.isnan(d)
    return keep


def edge_features(ca32, src, dst, dist):
    """Features of the ordered pairs (src=i -> dst=j) (`:233-273`), then the fp32 cast of
    `create_graphs.py:35`.  Returns (s [E,32] fp32, V [E,1,3] fp32)."""
    d = dist[src, dst]                                                       # fp64
    mu = np.linspace(0.0, RBF_DMAX, RBF_COUNT)                               # :233-236
    step = (RBF_DMAX - 0.0) / RBF_COUNT                                      # 1.25
    rbf = np.exp(-np.square((d[:, None] - mu[None, :]) / step))              # :237
    half = POSENC_COUNT // 2
    freqs = np.exp(2 * np.arange(half) * -(np.log(10000.0) / half))          # :380
    ang = (dst - src)[:, None] * freqs[None, :]                              # target idx - source idx :253
    pe = np.concatenate([np.cos(ang), np.sin(ang)], axis=-1)                 # :385
    diff = ca32[src] - ca32[dst]                                             # fp32, source - target :244
    sq = diff * diff
    nrm = np.sqrt((sq[:, 0] + sq[:, 1]) + sq[:, 2])[:, None]                 # np.linalg.norm in fp32
    direc = np.divide(diff, nrm, out=np.zeros_like(diff), where=nrm != 0)    # :360-365
    s = np.concatenate([rbf, pe], axis=-1).astype(np.float32)
    v = direc.astype(np.float32)[:, None, :]
    return s, v


def residue_graph(res_coords, edge_thresh=4.0, thresh_type="dist", keep_self_loops=True):
    """coords [n,4,3] fp32 (N, CA, C, O) -> (edge_index [2,E] int64 sorted by (src,dst), s, V).

    Composition of `compute_residue_edge_features(..., vectorize_features=True)` and `construct_graph`.
    """
    res_coords = np.asarray(res_coords, dtype=np.float32)
    ca = res_coords[:, 1, :]                                                 # :225
    dist = pairwise_ca_distance(ca)
    keep = select_edges(dist, edge_thresh, thresh_type, keep_self_loops)
    src, dst = np.nonzero(keep)                                              # row-major = (i asc, j asc)
    s, v = edge_features(ca, src, dst, dist)
    return np.stack([src, dst]).astype(np.int64), s, v


def _unit(x):
    """`normalize_vecs` (`:360-365`): x / ||x|| with 0 where the norm is 0."""
    nrm = np.linalg.norm(x, axis=-1, keepdims=True)
    return np.divide(x, nrm, out=np.zeros_like(x), where=nrm != 0)


def node_geometry_features(res_coords):
    """Geometric node features (`:27-92`): 6 dihedral scalars [cos(phi,psi,omega), sin(...)] and 3 vectors
    (forward, backward, virtual side chain).  The 11 amino-acid property columns (`:95-109`) are table
    look-ups and are not restated here."""
    x = np.asarray(res_coords)
    bb = x[:, :3, :].reshape(-1, 3)
    u = _unit(bb[1:] - bb[:-1])
    u0, u1, u2 = u[2:], u[1:-1], u[:-2]
    n1 = _unit(np.cross(u1, u0))
    n2 = _unit(np.cross(u2, u1))
    cosang = np.clip(np.sum(n1 * n2, -1), -1.0, 1.0)
    ang = np.arccos(cosang) * np.sign(np.sum(n1 * u2, -1))
    ang = np.pad(ang, [1, 2], "constant").reshape(-1, 3)
    scal = np.concatenate([np.cos(ang), np.sin(ang)], -1)
    ca = x[:, 1, :]
    fwd = _unit(ca[1:] - ca[:-1])
    f = np.pad(fwd, [(0, 1), (0, 0)])
    b = np.pad(-fwd, [(1, 0), (0, 0)])
    nvec = _unit(x[:, 0, :] - ca)
    cvec = _unit(x[:, 2, :] - ca)
    bis = _unit(nvec + cvec)
    perp = _unit(np.cross(cvec, nvec))
    side = -bis * np.sqrt(1 / 3) - perp * np.sqrt(2 / 3)
    return scal.astype(np.float32), np.stack([f, b, side], 1).astype(np.float32)

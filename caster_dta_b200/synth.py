"""Seeded synthetic protein / ligand inputs of Davis / KIBA / BindingDB shape.

PDB / ColabFold structures and RDKit are not available offline, so throughput and parity are measured on
synthetic backbones whose SHAPES (residues per protein, edges per residue, atoms per ligand) follow the
statistics of the shipped datasets (`data/deepdta_data/*`, `pretrained_model_downstream/model_summary.txt`).
Everything here is host-side numpy; graph construction from the coordinates is done by the CUDA featurizer
(`caster_dta_b200.featurizer`), or by the oracle in CPU tests.
"""
import numpy as np

CA_STEP = 3.8          # Angstrom between consecutive C-alpha atoms
NODE_SCALARS = 17      # 6 dihedral sin/cos + 11 residue property columns (`dataset_kwargs.json`)
NODE_VECTORS = 3
NUM_RESIDUE_TYPES = 20
MOL_NODE_FEATS, MOL_EDGE_FEATS, MOL_NODE_TYPES, MOL_EDGE_TYPES = 41, 9, 11, 5


def protein_lengths(shape, count, rng):
    """Residue counts per protein.  davis: U[300,1000]; kiba: log-uniform-ish (median ~620, clip 2000);
    bindingdb: mean ~558 in [25,2000]."""
    if shape == "davis":
        return rng.integers(300, 1001, size=count)
    if shape == "kiba":
        return np.clip(np.exp(rng.normal(np.log(620.0), 0.55, size=count)), 215, 2000).astype(np.int64)
    if shape == "bindingdb":
        return np.clip(np.exp(rng.normal(np.log(470.0), 0.6, size=count)), 25, 2000).astype(np.int64)
    if shape == "tiny":
        return rng.integers(20, 60, size=count)
    raise ValueError(shape)


def random_backbone(n, rng, self_avoiding=False, min_sep=4.5):
    """[n,4,3] fp32 (N, CA, C, O): C-alpha random walk with 3.8 A steps; the other backbone atoms are the
    C-alpha plus small Gaussian offsets.  `self_avoiding` rejects steps that land within `min_sep` of any
    non-adjacent residue, which reproduces the ~3 edges/residue of real 4 A radius graphs."""
    ca = np.zeros((n, 3), dtype=np.float64)
    for i in range(1, n):
        for _ in range(64):
            step = rng.normal(size=3)
            cand = ca[i - 1] + CA_STEP * step / np.linalg.norm(step)
            if not self_avoiding or i < 2:
                break
            if np.min(np.linalg.norm(ca[: i - 1] - cand, axis=1)) >= min_sep:
                break
        ca[i] = cand
    out = np.empty((n, 4, 3), dtype=np.float64)
    out[:, 1] = ca
    out[:, 0] = ca + rng.normal(scale=0.8, size=(n, 3))
    out[:, 2] = ca + rng.normal(scale=0.8, size=(n, 3))
    out[:, 3] = ca + rng.normal(scale=1.2, size=(n, 3))
    return out.astype(np.float32)


def random_backbone_fast(n, rng):
    """Vectorised `random_backbone(n, rng, self_avoiding=False)`: the same chain model (3.8 A C-alpha steps in uniformly
    random directions, Gaussian offsets for N / C / O) drawn with array calls -- a different random stream, so the loader's
    data sets use this one and the committed fixtures keep the step-by-step generator."""
    step = rng.normal(size=(n, 3))
    step /= np.linalg.norm(step, axis=1, keepdims=True)
    step[0] = 0.0
    ca = np.cumsum(CA_STEP * step, axis=0)
    out = np.empty((n, 4, 3), dtype=np.float64)
    out[:, 1] = ca
    out[:, 0] = ca + rng.normal(scale=0.8, size=(n, 3))
    out[:, 2] = ca + rng.normal(scale=0.8, size=(n, 3))
    out[:, 3] = ca + rng.normal(scale=1.2, size=(n, 3))
    return out.astype(np.float32)


def _safe_unit(x):
    nrm = np.sqrt((x * x).sum(-1, keepdims=True))
    return np.where(nrm > 0, x / np.where(nrm > 0, nrm, 1), 0.0)


def node_features(coords, rng):
    """(s [n,17] fp32, V [n,3,3] fp32, ntypes [n] int64): backbone dihedral sin/cos, unit orientation vectors
    and a virtual side-chain direction computed from the synthetic coordinates, plus 11 property columns
    drawn per residue type from a fixed random table (stand-in for the amino-acid property tables)."""
    n = coords.shape[0]
    chain = coords[:, :3, :].reshape(-1, 3).astype(np.float64)
    bond = _safe_unit(np.diff(chain, axis=0))
    a, b, c = bond[:-2], bond[1:-1], bond[2:]
    na, nb = _safe_unit(np.cross(a, b)), _safe_unit(np.cross(b, c))
    tors = np.arccos(np.clip((na * nb).sum(-1), -1, 1)) * np.sign((nb * a).sum(-1))
    tors = np.concatenate([[0.0], tors, [0.0, 0.0]]).reshape(n, 3)
    ca = coords[:, 1].astype(np.float64)
    fwd = np.zeros((n, 3))
    fwd[:-1] = _safe_unit(ca[1:] - ca[:-1])
    bwd = np.zeros((n, 3))
    bwd[1:] = -fwd[:-1]
    to_n, to_c = _safe_unit(coords[:, 0] - ca), _safe_unit(coords[:, 2] - ca)
    side = -_safe_unit(to_n + to_c) * np.sqrt(1 / 3) - _safe_unit(np.cross(to_c, to_n)) * np.sqrt(2 / 3)
    ntypes = rng.integers(0, NUM_RESIDUE_TYPES, size=n)
    table = np.random.default_rng(1234).random((NUM_RESIDUE_TYPES, 11))
    s = np.concatenate([np.cos(tors), np.sin(tors), table[ntypes]], -1).astype(np.float32)
    v = np.stack([fwd, bwd, side], 1).astype(np.float32)
    return s, v, ntypes.astype(np.int64)


def random_molecule(rng, lo=20, hi=46):
    """Ligand graph: chain + a few ring-closing bonds + self loops (~3.2 directed edges / atom, as in
    `model_summary.txt`: 3 791 edges / 1 197 atoms).  Returns x[a,41], edge_index[2,e], eattr[e,9], ntypes, etypes."""
    a = int(rng.integers(lo, hi + 1))
    pairs = [(i, i + 1) for i in range(a - 1)]
    for _ in range(max(1, a // 12)):
        i = int(rng.integers(0, a - 5))
        pairs.append((i, i + 5))
    src = [p[0] for p in pairs] + [p[1] for p in pairs] + list(range(a))
    dst = [p[1] for p in pairs] + [p[0] for p in pairs] + list(range(a))
    order = np.lexsort((dst, src))
    ei = np.stack([np.asarray(src)[order], np.asarray(dst)[order]]).astype(np.int64)
    e = ei.shape[1]
    selfloop = ei[0] == ei[1]
    etypes = np.where(selfloop, 0, rng.integers(1, MOL_EDGE_TYPES, size=e)).astype(np.int64)
    x = rng.random((a, MOL_NODE_FEATS)).astype(np.float32)
    eattr = rng.integers(0, 2, size=(e, MOL_EDGE_FEATS)).astype(np.float32)
    ntypes = rng.integers(0, MOL_NODE_TYPES, size=a).astype(np.int64)
    return x, ei, eattr, ntypes, etypes


def collate_molecules(mols):
    """Concatenate ligand graphs the way PyG `Batch.from_data_list` does (node offsets, `batch` vector)."""
    xs, eis, eas, nts, ets, bs = [], [], [], [], [], []
    off = 0
    for k, (x, ei, ea, nt, et) in enumerate(mols):
        xs.append(x); eis.append(ei + off); eas.append(ea); nts.append(nt); ets.append(et)
        bs.append(np.full(x.shape[0], k, dtype=np.int64))
        off += x.shape[0]
    return dict(x=np.concatenate(xs), edge_index=np.concatenate(eis, 1), eattr=np.concatenate(eas),
                ntypes=np.concatenate(nts), etypes=np.concatenate(ets), batch=np.concatenate(bs))


def protein_batch_coords(shape, pairs, seed, self_avoiding=False):
    """Backbones + node features for `pairs` proteins.  Returns dict with coords [N,4,3], ptr [B+1],
    x_s [N,17], x_v [N,3,3], ntypes [N], batch [N]."""
    rng = np.random.default_rng(seed)
    lens = protein_lengths(shape, pairs, rng)
    coords, xs, xv, nt, batch = [], [], [], [], []
    for k, n in enumerate(lens):
        c = random_backbone(int(n), rng, self_avoiding)
        s, v, t = node_features(c, rng)
        coords.append(c); xs.append(s); xv.append(v); nt.append(t)
        batch.append(np.full(int(n), k, dtype=np.int64))
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    return dict(coords=np.concatenate(coords), ptr=ptr, x_s=np.concatenate(xs), x_v=np.concatenate(xv),
                ntypes=np.concatenate(nt), batch=np.concatenate(batch))


def molecule_batch(pairs, seed, lo=20, hi=46):
    rng = np.random.default_rng(seed + 7919)
    return collate_molecules([random_molecule(rng, lo, hi) for _ in range(pairs)])


def conv_microbench_graph(num_edges, k=30, locality=40, seed=9):
    """Config-5 graph: N = E/k nodes, every node is the SOURCE of k edges whose targets are random distinct
    residues within +-`locality` positions (a synthetic chain kNN), sorted by (src, dst)."""
    rng = np.random.default_rng(seed)
    n = num_edges // k
    src = np.repeat(np.arange(n, dtype=np.int64), k)
    offs = np.stack([rng.permutation(2 * locality + 1)[:k] - locality for _ in range(min(n, 4096))])
    offs = offs[rng.integers(0, offs.shape[0], size=n)]
    dst = np.clip(src.reshape(n, k) + offs, 0, n - 1)
    dst.sort(axis=1)
    return np.stack([src, dst.reshape(-1)]).astype(np.int64), n

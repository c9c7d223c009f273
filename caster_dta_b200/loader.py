"""Pair data set, size-capped loader and fixed-shape ("bucketed") batches for the training / inference drivers.

Host-side mirror of the reference's data path for one step (`dataset/dual_dataset.py`):

    ProteinMoleculeDataset  (:88-125, unique-graph store + pair list)   -> SyntheticPairDataset
    PMD_DataLoader          (:399-421, batch sampler + collator)        -> PairBatchLoader
    PMD_BatchSampler        (:424-522)                                  -> batching.SizeCappedBatchSampler
    PMDCollator / Batch.from_data_list (:525-547)                       -> collate_pairs (+ device side:
                                                                           featurizer.protein_graph_batch)

What differs, B200-first: a protein is stored as its backbone COORDINATES (48 B / residue) instead of a featurized
graph (~4.3 KB / residue at kNN-30).  A batch crosses PCIe as ~1 MB of coordinates and the residue graph (node
features, kNN edges, RBF / positional / direction edge features, `utils/create_protein_features.py`) is rebuilt on the
device by `protein_graph_batch` inside the step -- 0.3 ms of kernels instead of a 115 MB host->device copy.

Fixed shapes for CUDA-graph replay: a real loader yields a different (residues, edges, atoms) triple every step.
`pad_pairs` appends DUMMY pairs -- real-looking filler proteins / ligands whose loss weight is 0 -- until the batch has
`slots` pairs, `n_pad` residues, `a_pad` atoms and `me_pad` ligand edges, where the padded sizes are the next bucket
boundaries.  Every graph of a batch is an independent connected component (no edge crosses pairs, the cross attention
is per pair, nothing normalises over the batch), so the dummy pairs change no real pair's prediction and contribute
exactly zero gradient; one captured graph per bucket then serves every batch that falls into it.
"""
import numpy as np
import torch

from . import synth
from .batching import SizeCappedBatchSampler
from .featurizer import knn_edge_count

MIN_DUMMY_RESIDUES = 32          # >= k for kNN-30 graphs, so that E = k * N holds for the padded batch too


class SyntheticPairDataset:
    """Unique proteins (backbone coordinates + residue identities), unique ligands (graphs) and (protein, ligand,
    affinity) pairs, shaped like Davis / KIBA / BindingDB (`synth.protein_lengths`).  Mirrors the unique-graph store
    of `ProteinMoleculeDataset` (`dual_dataset.py:88-125`): `__getitem__` returns (protein id, ligand id, y)."""

    def __init__(self, shape, num_pairs, seed=9, num_proteins=None, num_ligands=None, edge_thresh=30, thresh_type="num",
                 keep_self_loops=True, self_avoiding=False):
        rng = np.random.default_rng(seed)
        self.shape, self.edge_thresh, self.thresh_type, self.keep_self_loops = shape, edge_thresh, thresh_type, keep_self_loops
        num_proteins = num_pairs if num_proteins is None else num_proteins
        num_ligands = num_pairs if num_ligands is None else num_ligands
        lens = synth.protein_lengths(shape, num_proteins, rng)
        self.proteins = []
        for n in lens:
            c = synth.random_backbone(int(n), rng, True) if self_avoiding else synth.random_backbone_fast(int(n), rng)
            self.proteins.append(dict(coords=c, idents=rng.integers(0, synth.NUM_RESIDUE_TYPES, size=int(n)).astype(np.int64)))
        self.ligands = [synth.random_molecule(rng) for _ in range(num_ligands)]
        if num_proteins == num_pairs and num_ligands == num_pairs:
            self.pairs = [(i, i) for i in range(num_pairs)]
        else:
            self.pairs = [(int(rng.integers(0, num_proteins)), int(rng.integers(0, num_ligands))) for _ in range(num_pairs)]
        self.y = rng.normal(size=(num_pairs,)).astype(np.float32)
        # amino-acid property table: the reference's tables (utils/protein_definitions.py) are not available on the GPU box;
        # a fixed random [20, 11] table stands in (a look-up either way)
        self.aa_table = np.random.default_rng(1234).random((synth.NUM_RESIDUE_TYPES, 11)).astype(np.float32)

    def __len__(self):
        return len(self.pairs)

    def __getitem__(self, i):
        p, m = self.pairs[i]
        return p, m, float(self.y[i])

    def protein_edges(self, p):
        n = self.proteins[p]["coords"].shape[0]
        if self.thresh_type == "dist":
            return 3 * n                   # estimate only (the sampler's cost); the real count comes from the featurizer
        return knn_edge_count([n], self.edge_thresh, self.thresh_type, self.keep_self_loops)

    def sizes(self):
        """Per pair: protein nodes / edges, ligand nodes / edges -- what `PMD_BatchSampler` reads from the graphs."""
        pn = [self.proteins[p]["coords"].shape[0] for p, _ in self.pairs]
        pe = [self.protein_edges(p) for p, _ in self.pairs]
        mn = [self.ligands[m][0].shape[0] for _, m in self.pairs]
        me = [self.ligands[m][1].shape[1] for _, m in self.pairs]
        return pn, pe, mn, me

    def sampler(self, max_num, max_bsize=None, shuffle=True, generator=None, **kw):
        pn, pe, mn, me = self.sizes()
        return SizeCappedBatchSampler(pn, pe, mn, me, max_num=max_num, max_bsize=max_bsize, shuffle=shuffle,
                                      generator=generator, **kw)


def collate_pairs(dataset, indices):
    """Host-side collation of the pairs `indices` (`PMDCollator.__call__`, `dual_dataset.py:536-547`): concatenated
    backbone coordinates + `ptr`, residue identities, the collated ligand graph and the affinity vector (numpy)."""
    prots = [dataset.proteins[dataset.pairs[i][0]] for i in indices]
    ligs = [dataset.ligands[dataset.pairs[i][1]] for i in indices]
    lens = [p["coords"].shape[0] for p in prots]
    mol = synth.collate_molecules(ligs)
    return dict(coords=np.concatenate([p["coords"] for p in prots]), idents=np.concatenate([p["idents"] for p in prots]),
                ptr=np.concatenate([[0], np.cumsum(lens)]).astype(np.int64), mol=mol,
                mol_ptr=np.concatenate([[0], np.cumsum([l[0].shape[0] for l in ligs])]).astype(np.int64),
                y=dataset.y[np.asarray(indices, dtype=np.int64)].astype(np.float32), indices=list(indices))


def round_up(x, g):
    return (int(x) + g - 1) // g * g


class BucketSpec:
    """Bucket boundaries of the padded batch shapes (granularities in residues / atoms / ligand edges)."""

    def __init__(self, slots, node_gran=1024, atom_gran=128, mol_edge_gran=512):
        self.slots, self.node_gran, self.atom_gran, self.mol_edge_gran = int(slots), node_gran, atom_gran, mol_edge_gran

    def padded_sizes(self, pairs, nodes, atoms, mol_edges):
        """(n_pad, a_pad, me_pad) for a batch of `pairs` real pairs: every dummy slot needs >= MIN_DUMMY_RESIDUES
        residues and one atom (with its self loop)."""
        d = self.slots - pairs
        if d < 1:
            raise ValueError(f"{pairs} pairs do not leave a dummy slot (slots={self.slots})")
        n_pad = round_up(nodes + d * MIN_DUMMY_RESIDUES, self.node_gran)
        a_pad = round_up(atoms + d, self.atom_gran)
        me_pad = round_up(mol_edges + 1, self.mol_edge_gran)
        return n_pad, a_pad, me_pad


_FILLER = {}


def _filler_backbone(n):
    """Deterministic filler backbone for dummy proteins (a seeded random walk, prefix of one long chain)."""
    have = _FILLER.get("coords")
    if have is None or have.shape[0] < n:
        rng = np.random.default_rng(424242)
        have = synth.random_backbone_fast(max(n, 4096), rng)
        _FILLER["coords"] = have
        _FILLER["idents"] = rng.integers(0, synth.NUM_RESIDUE_TYPES, size=have.shape[0]).astype(np.int64)
    return have[:n], _FILLER["idents"][:n]


def pad_pairs(hb, spec, pin=True, global_pairs=None):
    """Collated host batch -> fixed-shape tensors (pinned by default) for the bucket it falls into.

    Dummy pairs fill the `spec.slots - pairs` free slots: each gets MIN_DUMMY_RESIDUES filler residues and one atom,
    the last one also takes the remaining residues / atoms / ligand edges up to the bucket boundary.  Their loss weight
    is 0; real pairs weigh 1 / pairs (mean-squared error over the real pairs, `train_model.py:565`) -- 1 / `global_pairs`
    when this is one rank's shard of a larger batch, so that the SUM of the ranks' gradients is the global-batch gradient.
    Returns (tensors, meta): `meta` holds the python-side shape facts (bucket key, hints for the device code)."""
    pairs = len(hb["ptr"]) - 1
    n, a, me = int(hb["ptr"][-1]), int(hb["mol"]["x"].shape[0]), int(hb["mol"]["edge_index"].shape[1])
    n_pad, a_pad, me_pad = spec.padded_sizes(pairs, n, a, me)
    d = spec.slots - pairs
    dummy_len = [MIN_DUMMY_RESIDUES] * d
    dummy_len[-1] += n_pad - n - d * MIN_DUMMY_RESIDUES
    dummy_atoms = [1] * d
    dummy_atoms[-1] += a_pad - a - d
    coords = np.empty((n_pad, 4, 3), np.float32)
    idents = np.empty((n_pad,), np.int64)
    coords[:n], idents[:n] = hb["coords"], hb["idents"]
    off = n
    for ln in dummy_len:
        c, t = _filler_backbone(ln)
        coords[off:off + ln], idents[off:off + ln] = c, t
        off += ln
    lens = np.concatenate([np.diff(hb["ptr"]), np.asarray(dummy_len, np.int64)])
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    # ligands: dummy atoms carry zero features; the spare edge slots are self loops of the last dummy atom
    mol = hb["mol"]
    x = np.zeros((a_pad, mol["x"].shape[1]), np.float32); x[:a] = mol["x"]
    nt = np.zeros((a_pad,), np.int64); nt[:a] = mol["ntypes"]
    mb = np.empty((a_pad,), np.int64); mb[:a] = mol["batch"]
    o = a
    for k, da in enumerate(dummy_atoms):
        mb[o:o + da] = pairs + k
        o += da
    ei = np.empty((2, me_pad), np.int64); ei[:, :me] = mol["edge_index"]
    ei[:, me:] = a_pad - 1
    ea = np.zeros((me_pad, mol["eattr"].shape[1]), np.float32); ea[:me] = mol["eattr"]
    et = np.zeros((me_pad,), np.int64); et[:me] = mol["etypes"]
    y = np.zeros((spec.slots,), np.float32); y[:pairs] = hb["y"]
    w = np.zeros((spec.slots,), np.float32); w[:pairs] = 1.0 / (pairs if global_pairs is None else global_pairs)
    t = dict(coords=coords, idents=idents, ptr=ptr, m_x=x, m_ei=ei, m_ea=ea, m_nt=nt, m_et=et, m_batch=mb, y=y, w=w)
    t = {k: torch.from_numpy(v) for k, v in t.items()}
    if pin:
        t = {k: v.pin_memory() for k, v in t.items()}
    meta = dict(pairs=pairs, nodes=n, atoms=a, mol_edges=me, n_pad=n_pad, a_pad=a_pad, me_pad=me_pad, slots=spec.slots,
                max_len=int(lens.max()), max_atoms=int(max(np.diff(hb["mol_ptr"]).max(), max(dummy_atoms))),
                lengths=[int(v) for v in lens])
    return t, meta


class PairBatchLoader:
    """Iterates (tensors, meta) fixed-shape pinned host batches: sampler -> collate_pairs -> pad_pairs
    (`PMD_DataLoader`, `dual_dataset.py:399-421`).  With `world_size > 1` every rank draws the SAME global batches
    (`max_bsize * world_size` pairs, cap `max_num * world_size`) and keeps the shard `parallel.shard_by_cost` assigns
    it -- equal pair counts, near-equal edge totals -- so that the step time (the max over ranks) is not set by one
    rank's oversized draw."""

    def __init__(self, dataset, max_num, max_bsize, spec=None, shuffle=True, seed=9, rank=0, world_size=1, pin=True, **kw):
        self.dataset, self.rank, self.world, self.pin = dataset, rank, world_size, pin
        self.max_bsize = max_bsize
        self.spec = spec or BucketSpec(max_bsize + 1)
        gen = torch.Generator().manual_seed(seed)
        self.sampler = dataset.sampler(max_num * world_size, max_bsize * world_size, shuffle, gen, **kw)
        self._pe = dataset.sizes()[1]

    def shard(self, indices):
        if self.world == 1:
            return list(indices)
        from .parallel import shard_by_cost
        parts = shard_by_cost([self._pe[i] for i in indices], self.world, equal_counts=True)
        return [indices[j] for j in parts[self.rank]]

    def __iter__(self):
        for indices in self.sampler:
            mine = self.shard(indices)
            if not mine:
                continue
            hb = collate_pairs(self.dataset, mine)
            yield pad_pairs(hb, self.spec, self.pin, global_pairs=len(indices))

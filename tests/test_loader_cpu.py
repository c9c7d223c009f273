"""Host-side data path of the training driver (CPU): pair data set, size-capped loader, bucket padding, rank sharding,
flat-buffer Adam -- and, on the oracle in fp64, the property the padding relies on: dummy pairs are inert."""
import numpy as np
import pytest
import torch

from caster_dta_b200 import loader, parallel, training
from caster_dta_b200.configs import caster_dta_2_2
from caster_dta_b200.featurizer import knn_edge_count


def test_knn_edge_count_matches_the_oracle_featurizer():
    from caster_dta_b200 import synth
    from oracle import featurizer_oracle
    rng = np.random.default_rng(1)
    for n, k, tt, ks in ((50, 30, "num", True), (20, 30, "num", True), (33, 10, "num", False), (41, 0.15, "prop", True)):
        ei, _, _ = featurizer_oracle.residue_graph(synth.random_backbone(n, rng), k, tt, ks)
        assert knn_edge_count([n], k, tt, ks) == ei.shape[1]
    with pytest.raises(ValueError):
        knn_edge_count([10], 4.0, "dist")


def test_pad_pairs_shapes_weights_and_bucket_key():
    ds = loader.SyntheticPairDataset("tiny", 20, seed=3, edge_thresh=10)
    spec = loader.BucketSpec(9, node_gran=256, atom_gran=64, mol_edge_gran=256)
    ld = loader.PairBatchLoader(ds, max_num=10_000_000, max_bsize=8, spec=spec, shuffle=False, pin=False)
    seen = 0
    for t, m in ld:
        seen += m["pairs"]
        assert t["coords"].shape == (m["n_pad"], 4, 3) and t["idents"].shape == (m["n_pad"],)
        assert t["ptr"].tolist()[0] == 0 and t["ptr"].tolist()[-1] == m["n_pad"] and len(t["ptr"]) == 10
        lens = np.diff(t["ptr"].numpy())
        assert (lens[m["pairs"]:] >= loader.MIN_DUMMY_RESIDUES).all()
        assert m["n_pad"] % 256 == 0 and m["a_pad"] % 64 == 0 and m["me_pad"] % 256 == 0
        assert t["m_x"].shape[0] == m["a_pad"] == t["m_batch"].shape[0] and t["m_ei"].shape == (2, m["me_pad"])
        assert int(t["m_ei"].max()) < m["a_pad"] and int(t["m_batch"].max()) == 8
        assert torch.equal(torch.unique(t["m_batch"]), torch.arange(9)), "every slot owns at least one atom"
        w = t["w"].numpy()
        assert np.allclose(w[:m["pairs"]], 1.0 / m["pairs"]) and (w[m["pairs"]:] == 0).all() and abs(w.sum() - 1) < 1e-6
        key = training.bucket_key(m, 10, "num")
        assert key == (9, m["n_pad"], 10 * m["n_pad"], m["a_pad"], m["me_pad"])      # E = k N once every graph has >= k nodes
    assert seen == 20


def test_loader_batches_follow_the_reference_sampler_rule():
    """Same batches as SizeCappedBatchSampler on the data set's size lists (itself pinned on the reference's sampler)."""
    ds = loader.SyntheticPairDataset("tiny", 30, seed=5, edge_thresh=10)
    pn, pe, mn, me = ds.sizes()
    from caster_dta_b200.batching import SizeCappedBatchSampler
    ref = [list(b) for b in SizeCappedBatchSampler(pn, pe, mn, me, max_num=6000, max_bsize=8, shuffle=False)]
    ld = loader.PairBatchLoader(ds, max_num=6000, max_bsize=8, shuffle=False, pin=False)
    got = [list(b) for b in ld.sampler]
    assert got == ref and sum(len(b) for b in got) == 30 and len(got) > 4


def test_shard_by_cost_equal_counts_and_balance():
    rng = np.random.default_rng(0)
    costs = [int(c) for c in rng.integers(9000, 30000, size=64)]
    for world in (2, 4, 8):
        parts = parallel.shard_by_cost(costs, world, equal_counts=True)
        assert sorted(i for p in parts for i in p) == list(range(64))
        assert {len(p) for p in parts} == {64 // world}
        loads = [sum(costs[i] for i in p) for p in parts]
        assert (max(loads) - min(loads)) / np.mean(loads) < 0.03
        naive = [sum(costs[i] for i in range(r, 64, world)) for r in range(world)]
        assert max(loads) <= max(naive)
    parts = parallel.shard_by_cost([5, 1, 1, 1, 1, 1], 2)
    assert sorted(sum(([5, 1, 1, 1, 1, 1][i] for i in p)) for p in parts) == [5, 5]


def test_world_size_two_loaders_partition_each_global_batch():
    ds = loader.SyntheticPairDataset("tiny", 32, seed=6, edge_thresh=10)
    lds = [loader.PairBatchLoader(ds, max_num=10_000_000, max_bsize=4, shuffle=True, seed=11, rank=r, world_size=2, pin=False)
           for r in range(2)]
    glob = list(loader.PairBatchLoader(ds, max_num=20_000_000, max_bsize=8, shuffle=True, seed=11, pin=False).sampler)
    for g, idx in zip(glob, zip(*[[ld.shard(b) for b in ld.sampler] for ld in lds])):
        assert sorted(idx[0] + idx[1]) == sorted(g) and len(idx[0]) == len(idx[1])


def test_flat_adam_equals_per_parameter_adam():
    torch.manual_seed(0)
    a = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    b = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    b.load_state_dict(a.state_dict())
    ref = torch.optim.Adam(a.parameters(), lr=1e-2)
    flat = parallel.FlatAdam(b, lr=1e-2, capturable=False)
    assert flat.numel == sum(p.numel() for p in b.parameters())
    x = torch.randn(11, 7)
    for _ in range(4):
        ref.zero_grad(set_to_none=True)
        a(x).square().sum().backward()
        ref.step()
        flat.reset()
        b(x).square().sum().backward()
        flat.sync()
        flat.step()
    for p, q in zip(a.parameters(), b.parameters()):
        assert torch.allclose(p, q, rtol=0, atol=1e-7)
    assert all(q.data_ptr() >= flat.flat_param.data_ptr() for q in b.parameters()), "parameters are views of the flat buffer"


def test_dummy_pairs_are_inert_in_the_oracle():
    """fp64 oracle: predictions of the real pairs, the loss and every parameter gradient are the same with and without
    the dummy pairs (and with a larger dummy), i.e. bucket padding does not change the training step."""
    from oracle import gvp_oracle, joint_oracle, pipeline
    import caster_dta_b200 as cg
    kw = caster_dta_2_2()
    torch.manual_seed(1)
    init = cg.JointGNN(kw["protein_gnn_kwargs"], kw["molecule_gnn_kwargs"], **kw["joint_gnn_kwargs"])
    ds = loader.SyntheticPairDataset("tiny", 5, seed=8, edge_thresh=10)
    hb = loader.collate_pairs(ds, [0, 1, 2, 3, 4])

    def run(t):
        p = {k: (v.detach().double().requires_grad_(v.numel() > 0) if v.dtype.is_floating_point else v)
             for k, v in init.state_dict().items()}
        prot = pipeline.featurize_batch(t["coords"].numpy(), t["ptr"].numpy(), t["idents"].numpy(), ds.aa_table, 10, "num",
                                        True, torch.float64)
        loss, pred = pipeline.train_loss(p, kw, prot, pipeline.molecule_dict(t, torch.float64), t["y"].double(), t["w"].double(),
                                         training=False)
        loss.backward()
        return float(loss), pred.detach()[:5], {k: v.grad for k, v in p.items() if torch.is_tensor(v) and v.grad is not None}

    # unpadded: the plain collated batch, weights 1/5
    n, a = int(hb["ptr"][-1]), hb["mol"]["x"].shape[0]
    plain = dict(coords=torch.from_numpy(hb["coords"]), ptr=torch.from_numpy(hb["ptr"]), idents=torch.from_numpy(hb["idents"]),
                 m_x=torch.from_numpy(hb["mol"]["x"]), m_ei=torch.from_numpy(hb["mol"]["edge_index"]),
                 m_ea=torch.from_numpy(hb["mol"]["eattr"]), m_nt=torch.from_numpy(hb["mol"]["ntypes"]),
                 m_et=torch.from_numpy(hb["mol"]["etypes"]), m_batch=torch.from_numpy(hb["mol"]["batch"]),
                 y=torch.from_numpy(hb["y"]), w=torch.full((5,), 0.2))
    base = run(plain)
    for slots, gran in ((6, 128), (8, 512)):
        t, m = loader.pad_pairs(hb, loader.BucketSpec(slots, node_gran=gran, atom_gran=32, mol_edge_gran=64), pin=False)
        assert m["n_pad"] > n and m["a_pad"] > a
        got = run(t)
        assert abs(got[0] - base[0]) <= 1e-12 * abs(base[0])
        assert float((got[1] - base[1]).abs().max()) <= 1e-12
        assert got[2].keys() == base[2].keys()
        for k in base[2]:
            assert float((got[2][k] - base[2][k]).abs().max()) <= 1e-12 * max(1.0, float(base[2][k].abs().max())), k

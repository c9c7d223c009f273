"""CPU oracle (TEST INFRASTRUCTURE) -- one training step of CASTER-DTA from backbone coordinates, composed from the
restatements in this package: per-protein featurization (`utils/create_protein_features.py:12-357`,
`utils/create_graphs.py:6-62`), `Batch.from_data_list` collation (`dataset/dual_dataset.py:543-544`), the model forward
(`models/joint_gnn.py:172-288`) and the MSE loss of `train_model.py:565`.  Only tests/, `__graft_entry__.smoke()` and the
reference / cpu_baseline legs of bench.py may use it.
"""
import numpy as np
import torch

from . import featurizer_oracle, gvp_oracle, joint_oracle


def featurize_batch(coords, ptr, idents, aa_table, edge_thresh=30, thresh_type="num", keep_self_loops=True, dtype=torch.float32):
    """Per-protein node + edge featurization and PyG-style collation.  coords [N,4,3] fp32, ptr [B+1], idents [N],
    aa_table [20, 11] -> the protein dict `joint_oracle.joint_forward` / `gvp_oracle.lba_encoder` take."""
    coords, ptr, idents = np.asarray(coords, np.float32), np.asarray(ptr, np.int64), np.asarray(idents, np.int64)
    aa_table = np.asarray(aa_table, np.float32)
    xs, xv, eis, ess, evs, batch = [], [], [], [], [], []
    for b in range(len(ptr) - 1):
        lo, hi = int(ptr[b]), int(ptr[b + 1])
        c = coords[lo:hi]
        geo_s, geo_v = featurizer_oracle.node_geometry_features(c)
        xs.append(np.concatenate([geo_s, aa_table[idents[lo:hi]]], -1).astype(np.float32))
        xv.append(geo_v)
        ei, es, ev = featurizer_oracle.residue_graph(c, edge_thresh, thresh_type, keep_self_loops)
        eis.append(ei + lo); ess.append(es); evs.append(ev)
        batch.append(np.full(hi - lo, b, np.int64))
    ei = torch.from_numpy(np.concatenate(eis, 1))
    return dict(x=(torch.from_numpy(np.concatenate(xs)).to(dtype), torch.from_numpy(np.concatenate(xv)).to(dtype)),
                edge_index=ei, ntypes=torch.from_numpy(idents.copy()), etypes=torch.zeros(ei.shape[1], dtype=torch.long),
                eattr=(torch.from_numpy(np.concatenate(ess)).to(dtype), torch.from_numpy(np.concatenate(evs)).to(dtype)),
                batch=torch.from_numpy(np.concatenate(batch)))


def molecule_dict(t, dtype=torch.float32):
    """The ligand half of a padded batch (`caster_dta_b200.loader.pad_pairs` tensors) as the oracle's molecule dict."""
    return dict(x=t["m_x"].to(dtype), edge_index=t["m_ei"], ntypes=t["m_nt"], etypes=t["m_et"], eattr=t["m_ea"].to(dtype),
                batch=t["m_batch"])


def train_loss(p, kw, prot, mol, y, w, gvp_masks=None, drop_masks=None, training=True):
    """Weighted squared error of the affinity predictions: sum_b w_b (pred_b - y_b)^2 (w = 1/pairs on real pairs gives
    `F.mse_loss`, `train_model.py:565`).  `gvp_masks[k] = ((ms0, mv0), (ms1, mv1))` per conv layer and `drop_masks`
    (site -> mask) inject the dropout masks recorded from the CUDA path."""
    pk = kw["protein_gnn_kwargs"]
    emb = gvp_oracle.lba_encoder(p, "protein_gnn.gnn_model.", prot["x"], prot["edge_index"], prot["ntypes"], prot["etypes"],
                                 prot["eattr"], pk["num_ntypes"], pk["num_etypes"], pk["num_convs"], pk["aggr"],
                                 drop_masks=gvp_masks)
    pred, _ = joint_oracle.joint_forward(p, kw, prot, mol, training=training, protein_embed=emb, drop_masks=drop_masks,
                                         num_graphs=int(y.shape[0]))
    err = pred.squeeze(-1) - y
    return (w * err * err).sum(), pred

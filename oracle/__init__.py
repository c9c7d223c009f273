"""CPU oracle for the CASTER-DTA GVP hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under `caster_dta_b200/` may import this package.  The only legitimate importers are
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs, and there
only as the checker or the CPU baseline, never as the thing shipped.

Contents (each function cites the reference file:line it restates; paths relative to the reference root):

* `gvp_oracle.py`        -- GVP / LayerNorm / Dropout / GVPConv / GVPConvLayer / LBA protein encoder
                            (`models/gvp_layers.py`, `models/protein_gnn.py`) as plain functional torch on CPU.
* `joint_oracle.py`      -- the rest of JointGNN (GINE molecule encoder, cross-attention, head;
                            `models/joint_gnn.py`, `models/molecule_gnn.py`) so that pairs/s can be timed on CPU.
* `featurizer_oracle.py` -- numpy restatement of `utils/create_protein_features.py:201-385` and
                            `utils/create_graphs.py:6-62` (edge set, RBF, positional encoding, directions).

Parity pinning: the reference ships NO tests or golden vectors for this path.  The oracle
is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF, run in the build container through the dependency
shim `tests/golden/ref_shim.py`; the generated vectors are committed under `tests/golden/*.npz` together with
`tests/golden/make_golden.py`.  `tests/test_oracle_golden.py` checks the oracle against those fixtures, and
(when `/root/reference` is present) against the live reference.
"""

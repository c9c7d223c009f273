"""castergvp: B200-native (sm_100a) GVP message passing for CASTER-DTA.

Public surface mirrors the reference's `models/gvp_layers.py`, `models/protein_gnn.py` (LBA encoder) and
`models/joint_gnn.py`; the arithmetic of the GVP stack and of the residue-graph featurizer runs in the C-ABI
library `libcastergvp.so` (see `include/castergvp.h`).  There is no CPU fallback.
"""
from .modules import (GVP, Dropout, GVPConv, GVPConvLayer, LayerNorm, _merge, _norm_no_nan, _split, randn,
                      tuple_cat, tuple_index, tuple_sum)
from .encoder import SelectableProteinModelWrapper, VectorProteinGNN_LBAModel
from .joint import JointGNN, load_state_dict_from_checkpoint
from . import batching
from .featurizer import residue_graph_batch, residue_node_features, aa_property_table, protein_graph_batch
from .ops import GraphPlan, gather_message_input, get_plan, segment_reduce

__all__ = ["GVP", "LayerNorm", "Dropout", "GVPConv", "GVPConvLayer", "tuple_sum", "tuple_cat", "tuple_index", "randn",
           "VectorProteinGNN_LBAModel", "SelectableProteinModelWrapper", "JointGNN", "load_state_dict_from_checkpoint",
           "residue_graph_batch", "residue_node_features", "aa_property_table", "protein_graph_batch", "GraphPlan", "get_plan", "gather_message_input", "segment_reduce"]

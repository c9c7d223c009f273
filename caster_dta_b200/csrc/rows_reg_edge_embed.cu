// Register-resident row program instance: gvp_edge + LayerNorm (models/protein_gnn.py:376) at the CASTER-DTA checkpoint dims
// (pretrained_model_downstream/model_kwargs.json).  See rows_reg.cuh.
#include "rows_reg.cuh"

using Spec = RowSpec<32, 1, 1, false, false, false, true, GvpC<33, 1, 32, 1, 1, CGVP_ACT_NONE, CGVP_ACT_NONE, 1>>;
CGVP_ROWS_INSTANCE(edge_embed, Spec)

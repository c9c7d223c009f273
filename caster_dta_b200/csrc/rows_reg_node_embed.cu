// Register-resident row program instance: gvp_node + LayerNorm (models/protein_gnn.py:375) at the CASTER-DTA checkpoint dims
// (pretrained_model_downstream/model_kwargs.json).  See rows_reg.cuh.
#include "rows_reg.cuh"

using Spec = RowSpec<17, 3, 20, false, false, false, true, GvpC<37, 3, 16, 4, 4, CGVP_ACT_NONE, CGVP_ACT_NONE, 1>>;
CGVP_ROWS_INSTANCE(node_embed, Spec)

// Register-resident row program instance: GVPConvLayer node update (models/gvp_layers.py:407-410) at the CASTER-DTA checkpoint dims
// (pretrained_model_downstream/model_kwargs.json).  See rows_reg.cuh.
#include "rows_reg.cuh"

using Spec = RowSpec<16, 4, 0, true, true, true, true, GvpC<16, 4, 64, 8, 8, CGVP_ACT_RELU, CGVP_ACT_NONE, 1>, GvpC<64, 8, 16, 4, 8, CGVP_ACT_NONE, CGVP_ACT_NONE, 1>>;
CGVP_ROWS_INSTANCE(node_update, Spec)

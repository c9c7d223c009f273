// Register-resident row program instance: gvp_norm_before_scalar + gvp_to_scalar (models/protein_gnn.py:385-386) at the CASTER-DTA checkpoint dims
// (pretrained_model_downstream/model_kwargs.json).  See rows_reg.cuh.
#include "rows_reg.cuh"

using Spec = RowSpec<16, 4, 0, false, true, false, false, GvpC<16, 4, 64, 0, 4, CGVP_ACT_RELU, CGVP_ACT_NONE, 1>>;
CGVP_ROWS_INSTANCE(readout, Spec)

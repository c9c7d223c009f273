// Dispatch of the register-resident fused row programs (rows_reg.cuh): one translation unit per compiled-in instance
// (rows_reg_node_embed.cu, rows_reg_edge_embed.cu, rows_reg_node_update.cu, rows_reg_readout.cu).
#include "cgvp_common.cuh"

bool cgvp_fast_paths_enabled();
bool rows_try_fwd_node_embed(const CgvpRowDesc*, const CgvpRowArgs*, cudaStream_t, int*);
bool rows_try_bwd_node_embed(const CgvpRowDesc*, const CgvpRowArgs*, const CgvpRowGradArgs*, void*, int64_t, cudaStream_t, int*);
int rows_pf_node_embed(const CgvpRowDesc*);
int rows_stashf_node_embed(const CgvpRowDesc*);
bool rows_try_fwd_edge_embed(const CgvpRowDesc*, const CgvpRowArgs*, cudaStream_t, int*);
bool rows_try_bwd_edge_embed(const CgvpRowDesc*, const CgvpRowArgs*, const CgvpRowGradArgs*, void*, int64_t, cudaStream_t, int*);
int rows_pf_edge_embed(const CgvpRowDesc*);
int rows_stashf_edge_embed(const CgvpRowDesc*);
bool rows_try_fwd_node_update(const CgvpRowDesc*, const CgvpRowArgs*, cudaStream_t, int*);
bool rows_try_bwd_node_update(const CgvpRowDesc*, const CgvpRowArgs*, const CgvpRowGradArgs*, void*, int64_t, cudaStream_t, int*);
int rows_pf_node_update(const CgvpRowDesc*);
int rows_stashf_node_update(const CgvpRowDesc*);
bool rows_try_fwd_readout(const CgvpRowDesc*, const CgvpRowArgs*, cudaStream_t, int*);
bool rows_try_bwd_readout(const CgvpRowDesc*, const CgvpRowArgs*, const CgvpRowGradArgs*, void*, int64_t, cudaStream_t, int*);
int rows_pf_readout(const CgvpRowDesc*);
int rows_stashf_readout(const CgvpRowDesc*);

// Returns 1 if a specialised kernel served the call (*rc = its result), 0 if the generic path must run.
int rows_fwd_special(const CgvpRowDesc* desc, const CgvpRowArgs* args, cudaStream_t st, int* rc) {
    if (!cgvp_fast_paths_enabled() || args->rows <= 0) return 0;
    return rows_try_fwd_node_embed(desc, args, st, rc) || rows_try_fwd_edge_embed(desc, args, st, rc) || rows_try_fwd_node_update(desc, args, st, rc) || rows_try_fwd_readout(desc, args, st, rc);
}
int rows_bwd_special(const CgvpRowDesc* desc, const CgvpRowArgs* args, const CgvpRowGradArgs* grads, void* ws, int64_t ws_bytes,
                     cudaStream_t st, int* rc) {
    if (!cgvp_fast_paths_enabled() || args->rows <= 0) return 0;
    return rows_try_bwd_node_embed(desc, args, grads, ws, ws_bytes, st, rc) || rows_try_bwd_edge_embed(desc, args, grads, ws, ws_bytes, st, rc) || rows_try_bwd_node_update(desc, args, grads, ws, ws_bytes, st, rc) || rows_try_bwd_readout(desc, args, grads, ws, ws_bytes, st, rc);
}
int rows_special_stash_floats(const CgvpRowDesc* desc) {
    if (!cgvp_fast_paths_enabled()) return 0;
    int f = 0;
    f = imax(f, rows_stashf_node_embed(desc)); f = imax(f, rows_stashf_edge_embed(desc)); f = imax(f, rows_stashf_node_update(desc)); f = imax(f, rows_stashf_readout(desc));
    return f;
}
int rows_special_partial_floats(const CgvpRowDesc* desc) {
    int pf = 0;
    pf = imax(pf, rows_pf_node_embed(desc)); pf = imax(pf, rows_pf_edge_embed(desc)); pf = imax(pf, rows_pf_node_update(desc)); pf = imax(pf, rows_pf_readout(desc));
    return pf;
}

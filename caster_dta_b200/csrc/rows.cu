// Fused per-row program: [one-hot ; x] (+ residual) -> LayerNorm -> GVP chain -> (+ residual) -> LayerNorm.
// One thread owns one row; the whole program runs out of a shared-memory tile (see cgvp_common.cuh for layout).
// Covers gvp_node / gvp_edge (+LN) (protein_gnn.py:375-376), the GVPConvLayer node update (gvp_layers.py:407-410),
// gvp_norm_before_scalar + gvp_to_scalar (protein_gnn.py:385-386) and stand-alone GVP / LayerNorm modules.
#include "cgvp_tile.cuh"

using namespace cgvp;

int rows_fwd_special(const CgvpRowDesc* desc, const CgvpRowArgs* args, cudaStream_t st, int* rc);
int rows_bwd_special(const CgvpRowDesc* desc, const CgvpRowArgs* args, const CgvpRowGradArgs* grads, void* ws, int64_t ws_bytes,
                     cudaStream_t st, int* rc);
int rows_special_partial_floats(const CgvpRowDesc* desc);
int64_t rows_wide_workspace_bytes(const CgvpRowDesc* desc, int64_t rows);
int rows_fwd_wide(const CgvpRowDesc* desc, const CgvpRowArgs* args, void* ws, int64_t ws_bytes, cudaStream_t st, int* rc);

struct RowsK {
    int in_s, in_v, onehot, has_res, pre_norm, n_gvp, post_res, post_norm;
    int s0, so_last, vo_last;            // scalar width of stage 0 (onehot + in_s); output dims
    GvpP g[CGVP_MAX_CHAIN];
    ChainCols cc;
    int x0_s, x0_v, x0_vpc;              // program input after the residual add (pre-LayerNorm0)
    int fin_s, fin_v, fin_vpc;           // pre-LayerNorm1 tensor (chain output, or x1 + mask1 * chain output)
    int stat0, stat1;
    // backward only
    int gy_s, gy_v;                      // upstream gradient (kept for the LayerNorm1 parameter gradients)
    int dx2_s, dx2_v;                    // gradient wrt the pre-LayerNorm1 tensor
    int gs[CGVP_MAX_CHAIN + 1], gv[CGVP_MAX_CHAIN + 1];   // gradient wrt stage k buffers (plane pitch cc.vpc[k])
    int dg[CGVP_MAX_CHAIN], dvh[CGVP_MAX_CHAIN];
    int dx0_s, dx0_v;                    // gradient wrt x0
    DwPlan dw;
    int goff[CGVP_MAX_CHAIN];            // offset of GVP k's gradient block in the partial arena
    int ln_off;                          // offset of [ln0_w, ln0_b, ln1_w, ln1_b] in the partial arena
    int ln0_n, ln1_n;
    int partial_floats;
    // launch geometry
    int R, rp, ncols, w_smem, woff[CGVP_MAX_CHAIN], wfloats[CGVP_MAX_CHAIN], wtotal;
    const float* wp[CGVP_MAX_CHAIN];
    long long rows;
    int ntiles;
    CgvpRowArgs a;
    CgvpRowGradArgs ga;
    float* partial;
};

// ---- cooperative staging ------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_rows_s(float4* T, int rp, int rv, int col4, int col0, int width,
                                             const float* __restrict__ src, const int* ridx) {
    if (!src || width <= 0) return;
    for (int i = threadIdx.x; i < rv * width; i += blockDim.x) {
        const int r = i / width, c = i - r * width;
        tile_at(T, rp, r, col4, col0 + c) = __ldg(src + (long long)ridx[r] * width + c);
    }
}
__device__ __forceinline__ void stage_rows_v(float4* T, int rp, int rv, int col4, int vpc, int nv,
                                             const float* __restrict__ src, const int* ridx) {
    if (!src || nv <= 0) return;
    const int w = 3 * nv;
    for (int i = threadIdx.x; i < rv * w; i += blockDim.x) {
        const int r = i / w, j = i - r * w, ch = j / 3, p = j - 3 * ch;
        tile_at(T, rp, r, col4 + p * vpc, ch) = __ldg(src + (long long)ridx[r] * w + j);
    }
    const int pad = vpc * 4 - nv;
    for (int i = threadIdx.x; i < rv * 3 * pad; i += blockDim.x) {
        const int r = i / (3 * pad), j = i - r * 3 * pad, p = j / pad, ch = nv + (j - p * pad);
        tile_at(T, rp, r, col4 + p * vpc, ch) = 0.f;
    }
}

// program input: one-hot prefix, features, optional residual addend with dropout mask
__device__ __forceinline__ void stage_program_input(const RowsK& K, float4* T, int* ridx, long long row0, int rv,
                                                    int s_col, int v_col, int vpc) {
    for (int r = threadIdx.x; r < rv; r += blockDim.x)
        ridx[r] = K.a.in_index ? K.a.in_index[row0 + r] : (int)(row0 + r);
    __syncthreads();
    if (K.onehot > 0) {
        for (int i = threadIdx.x; i < rv * K.onehot; i += blockDim.x) {
            const int r = i / K.onehot, c = i - r * K.onehot;
            tile_at(T, K.rp, r, s_col, c) = ((int)K.a.types[ridx[r]] == c) ? 1.f : 0.f;
        }
    }
    stage_rows_s(T, K.rp, rv, s_col, K.onehot, K.in_s, K.a.in_s, ridx);
    stage_rows_v(T, K.rp, rv, v_col, vpc, K.in_v, K.a.in_v, ridx);
    if (K.has_res) {
        __syncthreads();
        for (int i = threadIdx.x; i < rv * K.in_s; i += blockDim.x) {
            const int r = i / K.in_s, c = i - r * K.in_s;
            const long long g = (row0 + r) * K.in_s + c;
            const float m = K.a.mask0_s ? __ldg(K.a.mask0_s + g) : 1.f;
            tile_at(T, K.rp, r, s_col, c) += m * __ldg(K.a.h_s + g);
        }
        const int w = 3 * K.in_v;
        for (int i = threadIdx.x; i < rv * w; i += blockDim.x) {
            const int r = i / w, j = i - r * w, ch = j / 3, p = j - 3 * ch;
            const float m = K.a.mask0_v ? __ldg(K.a.mask0_v + (row0 + r) * K.in_v + ch) : 1.f;
            tile_at(T, K.rp, r, v_col + p * vpc, ch) += m * __ldg(K.a.h_v + (row0 + r) * w + j);
        }
    }
}

template <bool WS>
__device__ __forceinline__ const float* weights_for(const RowsK& K, const float* wsm, int k) {
    if (WS) return wsm + K.woff[k];   // shared-memory resident (compile-time choice keeps these loads LDS)
    return K.wp[k];
}

// forward part shared by the forward kernel and the backward recomputation
template <bool SAVE, bool WS>
__device__ __forceinline__ void rows_forward_row(const RowsK& K, float4* T, const float* wsm, int r, long long row) {
    const int rp = K.rp;
    if (K.pre_norm)
        ln_fwd_row(T, rp, r, K.s0, K.in_v, K.x0_s, K.x0_v, K.x0_vpc, K.cc.s[0], K.cc.v[0], K.a.ln0_w, K.a.ln0_b,
                   SAVE ? K.stat0 : -1);
    for (int k = 0; k < K.n_gvp; ++k) gvp_fwd_row<SAVE>(K.g[k], weights_for<WS>(K, wsm, k), T, rp, r, stage_io(K.cc, k));
    if (K.post_res) {
        const int L = K.n_gvp;
        for (int c = 0; c < K.so_last; ++c) {
            const float m = K.a.mask1_s ? __ldg(K.a.mask1_s + row * K.so_last + c) : 1.f;
            tile_at(T, rp, r, K.fin_s, c) = tile_at(T, rp, r, K.cc.s[0], c) + m * tile_at(T, rp, r, K.cc.s[L], c);
        }
        for (int c = 0; c < K.vo_last; ++c) {
            const float m = K.a.mask1_v ? __ldg(K.a.mask1_v + row * K.vo_last + c) : 1.f;
            for (int p = 0; p < 3; ++p)
                tile_at(T, rp, r, K.fin_v + p * K.fin_vpc, c) =
                    tile_at(T, rp, r, K.cc.v[0] + p * K.cc.vpc[0], c) + m * tile_at(T, rp, r, K.cc.v[L] + p * K.cc.vpc[L], c);
        }
    }
}

__device__ __forceinline__ void load_weights(const RowsK& K, float* wsm, bool backward) {
    if (!K.w_smem) return;
    for (int k = 0; k < K.n_gvp; ++k) copy_f4(wsm + K.woff[k], K.wp[k], backward ? K.g[k].total_floats : K.g[k].fwd_floats);
}

template <bool WS>
__global__ void __launch_bounds__(CGVP_THREADS) rows_fwd_kernel(const __grid_constant__ RowsK K) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* wsm = reinterpret_cast<float*>(smem);
    int* ridx = reinterpret_cast<int*>(wsm + K.wtotal);
    float4* T = reinterpret_cast<float4*>(ridx + K.R);
    load_weights(K, wsm, false);
    const int rp = K.rp;
    for (int i = threadIdx.x; i < K.ncols * rp; i += blockDim.x) T[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    for (int t = blockIdx.x; t < K.ntiles; t += gridDim.x) {
        const long long row0 = (long long)t * K.R;
        const int rv = (int)min((long long)K.R, K.rows - row0);
        stage_program_input(K, T, ridx, row0, rv, K.x0_s, K.x0_v, K.x0_vpc);
        __syncthreads();
        if (threadIdx.x < rv) {
            const int r = threadIdx.x;
            rows_forward_row<false, WS>(K, T, wsm, r, row0 + r);
            if (K.post_norm)
                ln_fwd_row(T, rp, r, K.so_last, K.vo_last, K.fin_s, K.fin_v, K.fin_vpc, K.fin_s, K.fin_v, K.a.ln1_w,
                           K.a.ln1_b, -1);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < rv * K.so_last; i += blockDim.x) {
            const int r = i / K.so_last, c = i - r * K.so_last;
            K.a.out_s[(row0 + r) * K.so_last + c] = tile_at(T, rp, r, K.fin_s, c);
        }
        if (K.vo_last > 0) {
            const int w = 3 * K.vo_last;
            for (int i = threadIdx.x; i < rv * w; i += blockDim.x) {
                const int r = i / w, j = i - r * w, ch = j / 3, p = j - 3 * ch;
                K.a.out_v[(row0 + r) * w + j] = tile_at(T, rp, r, K.fin_v + p * K.fin_vpc, ch);
            }
        }
        __syncthreads();
    }
}

template <int NSLOT, bool WS>
__global__ void __launch_bounds__(CGVP_THREADS) rows_bwd_kernel(const __grid_constant__ RowsK K) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* wsm = reinterpret_cast<float*>(smem);
    int* ridx = reinterpret_cast<int*>(wsm + K.wtotal);
    float4* T = reinterpret_cast<float4*>(ridx + K.R);
    load_weights(K, wsm, true);
    const int rp = K.rp, L = K.n_gvp;
    for (int i = threadIdx.x; i < K.ncols * rp; i += blockDim.x) T[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    float* partial = K.partial + (long long)blockIdx.x * K.partial_floats;
    DwAcc<NSLOT> dwacc;
    dwacc.init();
    float lnacc[4][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
    for (int t = blockIdx.x; t < K.ntiles; t += gridDim.x) {
        const long long row0 = (long long)t * K.R;
        const int rv = (int)min((long long)K.R, K.rows - row0);
        stage_program_input(K, T, ridx, row0, rv, K.x0_s, K.x0_v, K.x0_vpc);
        // upstream gradient (zero padded)
        for (int i = threadIdx.x; i < rv * K.so_last; i += blockDim.x) {
            const int r = i / K.so_last, c = i - r * K.so_last;
            tile_at(T, rp, r, K.gy_s, c) = __ldg(K.ga.d_out_s + (row0 + r) * K.so_last + c);
        }
        {
            const int pad = ((K.so_last + 3) & ~3) - K.so_last;
            for (int i = threadIdx.x; i < rv * pad; i += blockDim.x) tile_at(T, rp, i / pad, K.gy_s, K.so_last + i % pad) = 0.f;
        }
        if (K.vo_last > 0) {
            const int w = 3 * K.vo_last, vpc = K.fin_vpc;
            for (int i = threadIdx.x; i < rv * w; i += blockDim.x) {
                const int r = i / w, j = i - r * w, ch = j / 3, p = j - 3 * ch;
                tile_at(T, rp, r, K.gy_v + p * vpc, ch) = __ldg(K.ga.d_out_v + (row0 + r) * w + j);
            }
            const int pad = vpc * 4 - K.vo_last;
            for (int i = threadIdx.x; i < rv * 3 * pad; i += blockDim.x) {
                const int r = i / (3 * pad), j = i - r * 3 * pad, p = j / pad;
                tile_at(T, rp, r, K.gy_v + p * vpc, K.vo_last + (j - p * pad)) = 0.f;
            }
        }
        __syncthreads();
        if (threadIdx.x < rv) {
            const int r = threadIdx.x;
            const long long row = row0 + r;
            rows_forward_row<true, WS>(K, T, wsm, r, row);
            // LayerNorm1 backward -> dx2
            if (K.post_norm) {
                ln_fwd_row(T, rp, r, K.so_last, K.vo_last, K.fin_s, K.fin_v, K.fin_vpc, K.dx2_s, K.dx2_v, K.a.ln1_w,
                           K.a.ln1_b, K.stat1);   // output discarded (dx2 is overwritten next); stats kept
                ln_bwd_row(T, rp, r, K.so_last, K.vo_last, K.fin_s, K.fin_v, K.fin_vpc, K.gy_s, K.gy_v, K.fin_vpc,
                           K.dx2_s, K.dx2_v, K.fin_vpc, K.a.ln1_w, K.stat1);
            }
            // gradient entering the chain output (stage L): dx2 (* mask1 on the residual branch)
            for (int c4 = 0; c4 < ((K.so_last + 3) >> 2); ++c4) T[(K.gs[L] + c4) * rp + r] = T[(K.dx2_s + c4) * rp + r];
            for (int c4 = 0; c4 < 3 * K.fin_vpc; ++c4) T[(K.gv[L] + c4) * rp + r] = T[(K.dx2_v + c4) * rp + r];
            if (K.post_res) {
                if (K.a.mask1_s)
                    for (int c = 0; c < K.so_last; ++c) tile_at(T, rp, r, K.gs[L], c) *= __ldg(K.a.mask1_s + row * K.so_last + c);
                if (K.a.mask1_v)
                    for (int c = 0; c < K.vo_last; ++c) {
                        const float m = __ldg(K.a.mask1_v + row * K.vo_last + c);
                        for (int p = 0; p < 3; ++p) tile_at(T, rp, r, K.gv[L] + p * K.fin_vpc, c) *= m;
                    }
            }
            for (int k = L - 1; k >= 0; --k) {
                GradIO d;
                d.gs_in = K.gs[k + 1]; d.gv_in = K.gv[k + 1]; d.gv_in_pc = K.cc.vpc[k + 1];
                d.gs_out = K.gs[k]; d.gv_out = K.gv[k]; d.gv_out_pc = K.cc.vpc[k];
                d.dg = K.dg[k]; d.dvh = K.dvh[k]; d.dvh_pc = K.cc.vhpc[k];
                gvp_bwd_row(K.g[k], weights_for<WS>(K, wsm, k), T, rp, r, stage_io(K.cc, k), d);
            }
            if (K.post_res && L > 0) {   // residual branch: dx1 += dx2
                for (int c = 0; c < K.so_last; ++c) tile_at(T, rp, r, K.gs[0], c) += tile_at(T, rp, r, K.dx2_s, c);
                for (int p = 0; p < 3; ++p)
                    for (int c = 0; c < K.vo_last; ++c)
                        tile_at(T, rp, r, K.gv[0] + p * K.cc.vpc[0], c) += tile_at(T, rp, r, K.dx2_v + p * K.fin_vpc, c);
            }
            if (K.pre_norm)
                ln_bwd_row(T, rp, r, K.s0, K.in_v, K.x0_s, K.x0_v, K.x0_vpc, K.gs[0], K.gv[0], K.cc.vpc[0], K.dx0_s,
                           K.dx0_v, K.x0_vpc, K.a.ln0_w, K.stat0);
        }
        __syncthreads();
        // ---- cooperative epilogue: input gradients, weight gradients, LayerNorm parameter gradients
        if (K.ga.d_in_s) {
            for (int i = threadIdx.x; i < rv * K.in_s; i += blockDim.x) {
                const int r = i / K.in_s, c = i - r * K.in_s;
                K.ga.d_in_s[(long long)ridx[r] * K.in_s + c] = tile_at(T, rp, r, K.dx0_s, K.onehot + c);
            }
        }
        if (K.ga.d_in_v && K.in_v > 0) {
            const int w = 3 * K.in_v;
            for (int i = threadIdx.x; i < rv * w; i += blockDim.x) {
                const int r = i / w, j = i - r * w, ch = j / 3, p = j - 3 * ch;
                K.ga.d_in_v[(long long)ridx[r] * w + j] = tile_at(T, rp, r, K.dx0_v + p * K.x0_vpc, ch);
            }
        }
        if (K.has_res && K.ga.d_h_s) {
            for (int i = threadIdx.x; i < rv * K.in_s; i += blockDim.x) {
                const int r = i / K.in_s, c = i - r * K.in_s;
                const long long g = (row0 + r) * K.in_s + c;
                K.ga.d_h_s[g] = tile_at(T, rp, r, K.dx0_s, c) * (K.a.mask0_s ? __ldg(K.a.mask0_s + g) : 1.f);
            }
            const int w = 3 * K.in_v;
            for (int i = threadIdx.x; i < rv * w; i += blockDim.x) {
                const int r = i / w, j = i - r * w, ch = j / 3, p = j - 3 * ch;
                const float m = K.a.mask0_v ? __ldg(K.a.mask0_v + (row0 + r) * K.in_v + ch) : 1.f;
                K.ga.d_h_v[(row0 + r) * w + j] = tile_at(T, rp, r, K.dx0_v + p * K.x0_vpc, ch) * m;
            }
        }
        dwacc.tile(K.dw, T, rp, rv, partial);
        // LayerNorm parameter gradients: channel c = threadIdx.x (+128)
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int c = threadIdx.x + s * CGVP_THREADS;
            if (K.pre_norm && c < K.s0) {
                float aw = 0.f, ab = 0.f;
                for (int r = 0; r < rv; ++r) {
                    const float4 st = T[K.stat0 * rp + r];
                    const float dy = tile_at(T, rp, r, K.gs[0], c);
                    aw += dy * (tile_at(T, rp, r, K.x0_s, c) - st.x) * st.y;
                    ab += dy;
                }
                lnacc[0][s] += aw; lnacc[1][s] += ab;
            }
            if (K.post_norm && c < K.so_last) {
                float aw = 0.f, ab = 0.f;
                for (int r = 0; r < rv; ++r) {
                    const float4 st = T[K.stat1 * rp + r];
                    const float dy = tile_at(T, rp, r, K.gy_s, c);
                    aw += dy * (tile_at(T, rp, r, K.fin_s, c) - st.x) * st.y;
                    ab += dy;
                }
                lnacc[2][s] += aw; lnacc[3][s] += ab;
            }
        }
        __syncthreads();
    }
    dwacc.flush(K.dw, partial);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int c = threadIdx.x + s * CGVP_THREADS;
        if (K.pre_norm && c < K.s0) { partial[K.ln_off + c] = lnacc[0][s]; partial[K.ln_off + K.ln0_n + c] = lnacc[1][s]; }
        if (K.post_norm && c < K.so_last) {
            partial[K.ln_off + 2 * K.ln0_n + c] = lnacc[2][s];
            partial[K.ln_off + 2 * K.ln0_n + K.ln1_n + c] = lnacc[3][s];
        }
    }
}

// ---- host side -----------------------------------------------------------------------------------------------------
static int build_rows_k(const CgvpRowDesc* desc, bool backward, RowsK& K) {
    memset(&K, 0, sizeof(K));
    CGVP_REQUIRE(desc, "rows: null descriptor");
    CGVP_REQUIRE(desc->n_gvp >= 0 && desc->n_gvp <= CGVP_MAX_CHAIN, "rows: n_gvp %d out of range", desc->n_gvp);
    CGVP_REQUIRE(desc->in_s > 0 && desc->in_v >= 0 && desc->onehot >= 0, "rows: bad input dims");
    CGVP_REQUIRE(!(desc->has_residual_in && desc->onehot), "rows: residual input and one-hot prefix are exclusive");
    K.in_s = desc->in_s; K.in_v = desc->in_v; K.onehot = desc->onehot; K.has_res = desc->has_residual_in;
    K.pre_norm = desc->pre_norm; K.n_gvp = desc->n_gvp; K.post_res = desc->post_residual; K.post_norm = desc->post_norm;
    K.s0 = K.onehot + K.in_s;
    int cs = K.s0, cv = K.in_v;
    for (int k = 0; k < K.n_gvp; ++k) {
        if (cgvp_validate_gvp(desc->gvp[k], "rows")) return -1;
        K.g[k] = make_gvp_p(desc->gvp[k]);
        CGVP_REQUIRE(K.g[k].si == cs && K.g[k].vi == cv, "rows: GVP %d expects (%d,%d) but receives (%d,%d)", k,
                     K.g[k].si, K.g[k].vi, cs, cv);
        cs = K.g[k].so; cv = K.g[k].vo;
    }
    K.so_last = cs; K.vo_last = cv;
    CGVP_REQUIRE(!K.post_res || (cs == K.s0 && cv == K.in_v && K.n_gvp > 0), "rows: post_residual needs matching dims");
    if (backward && (K.pre_norm || K.post_norm))
        CGVP_REQUIRE(K.s0 <= 2 * CGVP_THREADS && K.so_last <= 2 * CGVP_THREADS,
                     "rows: LayerNorm wider than %d is not supported in backward", 2 * CGVP_THREADS);
    int col = 0;
    const int s04 = cdiv(K.s0, 4), v04 = cdiv(K.in_v, 4);
    if (K.n_gvp > 0) {
        K.cc = plan_chain_cols(K.g, K.n_gvp, true, backward, col);
        col += K.cc.ncols;
    } else {
        K.cc.s[0] = col; col += s04; K.cc.v[0] = col; K.cc.vpc[0] = v04; col += 3 * v04;
    }
    const int L = K.n_gvp;
    // pre-LN0 input: forward may normalise in place; backward needs x0 preserved
    if (K.pre_norm && backward) { K.x0_s = col; col += s04; K.x0_v = col; K.x0_vpc = v04; col += 3 * v04; }
    else { K.x0_s = K.cc.s[0]; K.x0_v = K.cc.v[0]; K.x0_vpc = K.cc.vpc[0]; }
    const int so4 = cdiv(K.so_last, 4), vo4 = K.cc.vpc[L];
    if (K.post_res && backward) { K.fin_s = col; col += so4; K.fin_v = col; col += 3 * vo4; }
    else { K.fin_s = K.cc.s[L]; K.fin_v = K.cc.v[L]; }
    K.fin_vpc = vo4;
    K.stat0 = col++; K.stat1 = col++;
    if (backward) {
        K.gy_s = col; col += so4; K.gy_v = col; col += 3 * vo4;
        if (K.post_norm) { K.dx2_s = col; col += so4; K.dx2_v = col; col += 3 * vo4; }
        else { K.dx2_s = K.gy_s; K.dx2_v = K.gy_v; }
        for (int k = 0; k <= L; ++k) {   // gradient wrt stage k: [dS_in ; dvn] of GVP k, or the chain output
            K.gs[k] = col; col += k < L ? K.g[k].ksd4 : so4;
            K.gv[k] = col; col += 3 * K.cc.vpc[k];
        }
        for (int k = 0; k < L; ++k) { K.dg[k] = col; col += K.g[k].vo4; K.dvh[k] = col; col += 3 * K.g[k].h4; }
        if (K.pre_norm) { K.dx0_s = col; col += s04; K.dx0_v = col; col += 3 * v04; }
        else { K.dx0_s = K.gs[0]; K.dx0_v = K.gv[0]; }
        int goff = 0;
        for (int k = 0; k < L; ++k) {
            K.goff[k] = goff;
            GradIO d;
            d.gs_in = K.gs[k + 1]; d.gv_in = K.gv[k + 1]; d.gv_in_pc = K.cc.vpc[k + 1];
            d.gs_out = K.gs[k]; d.gv_out = K.gv[k]; d.gv_out_pc = K.cc.vpc[k];
            d.dg = K.dg[k]; d.dvh = K.dvh[k]; d.dvh_pc = K.cc.vhpc[k];
            dw_add_gvp(K.dw, K.g[k], stage_io(K.cc, k), d, goff);
            goff += K.g[k].fwd_floats;
        }
        K.ln_off = goff;
        K.ln0_n = K.pre_norm ? K.s0 : 0;
        K.ln1_n = K.post_norm ? K.so_last : 0;
        K.partial_floats = (int)align_up(goff + 2 * K.ln0_n + 2 * K.ln1_n, 4);
    }
    K.ncols = col;
    int wt = 0;
    for (int k = 0; k < L; ++k) {
        K.woff[k] = wt;
        K.wfloats[k] = backward ? K.g[k].total_floats : K.g[k].fwd_floats;
        wt += (int)align_up(K.wfloats[k], 4);
    }
    K.wtotal = wt;
    return 0;
}

// choose rows-per-tile and whether the weights live in shared memory
static int choose_geometry(RowsK& K, int smem_max, size_t* smem_bytes) {
    for (int pass = 0; pass < 2; ++pass) {
        const int wbytes = pass == 0 ? K.wtotal * 4 : 0;
        for (int R = CGVP_THREADS; R >= 8; R = R > 32 ? R - 32 : R / 2) {   // 128, 96, 64, 32, 16, 8
            const size_t need = (size_t)wbytes + (size_t)R * 4 + (size_t)K.ncols * (R + 1) * 16 + 16;
            if (need <= (size_t)smem_max) {
                K.R = R; K.rp = R + 1; K.w_smem = pass == 0;
                if (!K.w_smem) K.wtotal = 0;
                *smem_bytes = need;
                return 0;
            }
        }
    }
    cgvp_set_error("rows: tile of %d float4 columns does not fit in shared memory", K.ncols);
    return -1;
}

int rows_special_stash_floats(const CgvpRowDesc* desc);
extern "C" int64_t cgvp_rows_stash_floats(const CgvpRowDesc* desc) {
    if (!desc) return 0;
    return rows_special_stash_floats(desc);
}

extern "C" int64_t cgvp_rows_workspace_bytes(const CgvpRowDesc* desc, int64_t rows, int32_t backward) {
    if (!backward) return 16 + rows_wide_workspace_bytes(desc, rows);
    RowsK K;
    if (build_rows_k(desc, true, K)) return -1;
    const int sms = cgvp_num_sms() > 0 ? cgvp_num_sms() : 148;
    const int pf = imax(K.partial_floats, rows_special_partial_floats(desc));
    return (int64_t)(2 * sms + 1) * pf * 4 + 256;
}

static int check_args(const RowsK& K, const CgvpRowArgs* a) {
    CGVP_REQUIRE(a && a->rows >= 0, "rows: null args");
    CGVP_REQUIRE(a->in_s && (K.in_v == 0 || a->in_v), "rows: null input");
    CGVP_REQUIRE(!K.onehot || a->types, "rows: one-hot prefix needs types");
    CGVP_REQUIRE(!K.has_res || (a->h_s && (K.in_v == 0 || a->h_v)), "rows: residual addend missing");
    CGVP_REQUIRE(!K.pre_norm || (a->ln0_w && a->ln0_b), "rows: LayerNorm0 parameters missing");
    CGVP_REQUIRE(!K.post_norm || (a->ln1_w && a->ln1_b), "rows: LayerNorm1 parameters missing");
    CGVP_REQUIRE(K.n_gvp == 0 || a->h_packed, "rows: packed weights missing");
    for (int k = 0; k < K.n_gvp; ++k)
        CGVP_REQUIRE(a->h_packed[k] && ((uintptr_t)a->h_packed[k] & 15) == 0, "rows: packed block %d null/unaligned", k);
    return 0;
}

extern "C" int32_t cgvp_rows_fwd(const CgvpRowDesc* desc, const CgvpRowArgs* args, void* ws, int64_t ws_bytes,
                                 cgvp_stream_t stream) {
    RowsK K;
    if (build_rows_k(desc, false, K)) return -1;
    if (check_args(K, args)) return -1;
    CGVP_REQUIRE(args->out_s && (K.vo_last == 0 || args->out_v), "rows: null output");
    if (args->rows == 0) return 0;
    {   // register-resident specialised kernels for the dims they were compiled for (rows_reg.cu)
        int rc = 0;
        if (rows_fwd_wide(desc, args, ws, ws_bytes, (cudaStream_t)stream, &rc)) return rc;
        if (rows_fwd_special(desc, args, (cudaStream_t)stream, &rc)) return rc;
    }
    size_t smem = 0;
    if (choose_geometry(K, cgvp_max_smem_optin(), &smem)) return -1;
    K.a = *args;
    for (int k = 0; k < K.n_gvp; ++k) K.wp[k] = args->h_packed[k];
    K.rows = args->rows;
    K.ntiles = (int)cdiv64(args->rows, K.R);
    const int sms = cgvp_num_sms();
    int per_sm = (int)((size_t)cgvp_max_smem_optin() / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    const int grid = K.ntiles < sms * per_sm ? K.ntiles : sms * per_sm;
    cgvp_prof_begin(CGVP_K_ROWS_FWD, (cudaStream_t)stream);
    if (K.w_smem) {
        CGVP_CUDA(cudaFuncSetAttribute(rows_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rows_fwd_kernel<true><<<grid, CGVP_THREADS, smem, (cudaStream_t)stream>>>(K);
    } else {
        CGVP_CUDA(cudaFuncSetAttribute(rows_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rows_fwd_kernel<false><<<grid, CGVP_THREADS, smem, (cudaStream_t)stream>>>(K);
    }
    cgvp_prof_end(CGVP_K_ROWS_FWD, (cudaStream_t)stream);
    CGVP_LAUNCH_CHECK("rows_fwd_kernel");
    return 0;
}

extern "C" int32_t cgvp_rows_bwd(const CgvpRowDesc* desc, const CgvpRowArgs* args, const CgvpRowGradArgs* grads,
                                 void* ws, int64_t ws_bytes, cgvp_stream_t stream) {
    RowsK K;
    if (build_rows_k(desc, true, K)) return -1;
    if (check_args(K, args)) return -1;
    CGVP_REQUIRE(grads && grads->d_out_s && (K.vo_last == 0 || grads->d_out_v), "rows_bwd: null upstream gradient");
    CGVP_REQUIRE(K.n_gvp == 0 || grads->h_packed_grads, "rows_bwd: packed gradient blocks missing");
    CGVP_REQUIRE(!K.has_res || !grads->d_h_s || K.in_v == 0 || grads->d_h_v, "rows_bwd: d_h_v missing");
    cudaStream_t st = (cudaStream_t)stream;
    {
        int rc = 0;
        if (rows_bwd_special(desc, args, grads, ws, ws_bytes, st, &rc)) return rc;
    }
    size_t smem = 0;
    if (choose_geometry(K, cgvp_max_smem_optin(), &smem)) return -1;
    K.a = *args;
    K.ga = *grads;
    for (int k = 0; k < K.n_gvp; ++k) K.wp[k] = args->h_packed[k];
    K.rows = args->rows;
    K.ntiles = (int)cdiv64(args->rows, K.R);
    const int sms = cgvp_num_sms();
    int per_sm = (int)((size_t)cgvp_max_smem_optin() / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2) per_sm = 2;
    int grid = K.ntiles < sms * per_sm ? K.ntiles : sms * per_sm;
    if (grid < 1) grid = 1;
    const int64_t need = (int64_t)(grid + 1) * K.partial_floats * 4;
    CGVP_REQUIRE(ws && ws_bytes >= need && ((uintptr_t)ws & 15) == 0, "rows_bwd: workspace too small (%lld < %lld)",
                 (long long)ws_bytes, (long long)need);
    K.partial = reinterpret_cast<float*>(ws);
    float* reduced = K.partial + (int64_t)grid * K.partial_floats;
    if (K.partial_floats > 0) CGVP_CUDA(cudaMemsetAsync(ws, 0, (size_t)need, st));
    if (args->rows > 0) {
        // two register-resident 4x4 weight-gradient blocks per thread; larger chains spill to the global partial
        cgvp_prof_begin(CGVP_K_ROWS_BWD, st);
        if (K.w_smem) {
            CGVP_CUDA(cudaFuncSetAttribute(rows_bwd_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            rows_bwd_kernel<2, true><<<grid, CGVP_THREADS, smem, st>>>(K);
        } else {
            CGVP_CUDA(cudaFuncSetAttribute(rows_bwd_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            rows_bwd_kernel<2, false><<<grid, CGVP_THREADS, smem, st>>>(K);
        }
        cgvp_prof_end(CGVP_K_ROWS_BWD, st);
        CGVP_LAUNCH_CHECK("rows_bwd_kernel");
    }
    if (K.partial_floats > 0) {
        CgvpSeg seg[CGVP_MAX_SEGS];
        memset(seg, 0, sizeof(seg));
        int ns = 0;
        for (int k = 0; k < K.n_gvp; ++k) { seg[ns].dst = grads->h_packed_grads[k]; seg[ns].off = K.goff[k]; seg[ns].n = K.g[k].fwd_floats; ++ns; }
        if (K.pre_norm) {
            seg[ns].dst = grads->d_ln0_w; seg[ns].off = K.ln_off; seg[ns].n = K.ln0_n; ++ns;
            seg[ns].dst = grads->d_ln0_b; seg[ns].off = K.ln_off + K.ln0_n; seg[ns].n = K.ln0_n; ++ns;
        }
        if (K.post_norm) {
            seg[ns].dst = grads->d_ln1_w; seg[ns].off = K.ln_off + 2 * K.ln0_n; seg[ns].n = K.ln1_n; ++ns;
            seg[ns].dst = grads->d_ln1_b; seg[ns].off = K.ln_off + 2 * K.ln0_n + K.ln1_n; seg[ns].n = K.ln1_n; ++ns;
        }
        const int rc = cgvp_reduce_partials(K.partial, grid, K.partial_floats, reduced, seg, ns, st);
        if (rc) return rc;
    }
    return 0;
}

// Register-resident fused row programs for compile-time known dims (the CASTER-DTA checkpoint encoder):
//     x = [onehot(types) ; in_s], in_v  (+ mask0 * h)  -> LayerNorm0 -> GVP chain -> (+ residual) -> LayerNorm1
// Same contract as rows.cu (cgvp_rows_fwd / cgvp_rows_bwd), but one thread keeps its whole row in registers
// (cgvp_reg.cuh) and one warp = 32 consecutive rows runs independently (cgvp_warp.cuh).  Instances:
//     gvp_node + LN, gvp_edge + LN              models/protein_gnn.py:375-376
//     GVPConvLayer node update                   models/gvp_layers.py:407-410
//     gvp_norm_before_scalar + gvp_to_scalar     models/protein_gnn.py:385-386
#pragma once
#include "cgvp_warp.cuh"

using namespace cgvpr;

struct NoG {
    static constexpr int SI = 0, VI = 0, SO = 0, VO = 0, FWD_FLOATS = 0, TOTAL_FLOATS = 0;
};

struct RowsRegArgs {
    CgvpRowArgs a;
    CgvpRowGradArgs g;
    const float* wp[2];
    float* partial;
    int ntiles;
};

template <int IN_S_, int IN_V_, int ONEHOT_, bool RES_IN_, bool PRE_NORM_, bool POST_RES_, bool POST_NORM_, class G0_, class G1_ = NoG>
struct RowSpec {
    static constexpr int IN_S = IN_S_, IN_V = IN_V_, ONEHOT = ONEHOT_;
    static constexpr bool RES_IN = RES_IN_, PRE_NORM = PRE_NORM_, POST_RES = POST_RES_, POST_NORM = POST_NORM_;
    using G0 = G0_; using G1 = G1_;
    static constexpr int NG = std::is_same<G1_, NoG>::value ? 1 : 2;
    static constexpr int S0 = ONEHOT + IN_S, V0 = IN_V, V01 = max1(V0);
    static constexpr int OS = NG == 2 ? G1::SO : G0::SO, OV = NG == 2 ? G1::VO : G0::VO, OV1 = max1(OV);
    static_assert(G0::SI == S0 && G0::VI == V0, "row program: first GVP input dims");
    static_assert(NG == 1 || (G1::SI == G0::SO && G1::VI == G0::VO), "row program: chain dims");
    static_assert(!POST_RES || (OS == S0 && OV == V0), "row program: residual dims");
    static_assert(!(RES_IN && ONEHOT > 0), "row program: residual input and one-hot prefix are exclusive");
    // shared memory: weights, then LayerNorm parameters
    static constexpr int WF1 = G0::FWD_FLOATS, WF = WF1 + G1::FWD_FLOATS;
    static constexpr int WT1 = G0::TOTAL_FLOATS, WT = WT1 + G1::TOTAL_FLOATS;
    static constexpr int LNP = pad4(2 * S0) + pad4(2 * OS);          // [ln0_w ; ln0_b] [ln1_w ; ln1_b]
    // gradient arena: GVP blocks, then one [4][pad4(2S)] block per LayerNorm (row 0 = [d_w ; d_b])
    static constexpr int GO0 = 0, GO1 = G0::FWD_FLOATS, LN0_OFF = GO1 + G1::FWD_FLOATS;
    static constexpr int LN1_OFF = LN0_OFF + (PRE_NORM ? 4 * pad4(2 * S0) : 0);
    static constexpr int PF = LN1_OFF + (POST_NORM ? 4 * pad4(2 * OS) : 0);
    static constexpr int sink_cols_g1() { if constexpr (NG == 2) return sink_cols<G1>(); else return 0; }
    static constexpr int STG_COLS = imax(imax(sink_cols<G0>(), sink_cols_g1()),
                                         imax(PRE_NORM ? 1 + cdiv4(2 * S0) : 0, POST_NORM ? 1 + cdiv4(2 * OS) : 0));
    static constexpr int STG_FLOATS = STG_COLS * CGVP_WPITCH * 4;
    static constexpr int PER_WARP = PF + STG_FLOATS;
    // training stash row: s' (pre-activation scalars) of GVP 0 [and GVP 1]
    static constexpr int g1_so() { if constexpr (NG == 2) return G1::SO; else return 0; }
    static constexpr int STASHF = pad4(G0::SO) + pad4(g1_so());
    static constexpr int bwd_warps() {
        for (int w = 16; w >= 4; w -= 4)
            if ((size_t)(WT + LNP + w * PER_WARP) * 4 + 1024 <= (size_t)CGVP_SMEM_OPTIN) return w;
        return 2;
    }
    static constexpr int BW = bwd_warps();
    static constexpr size_t smem_fwd() { return (size_t)(WF + LNP) * 4; }
    static constexpr size_t smem_bwd() { return (size_t)(WT + LNP + BW * PER_WARP) * 4; }
    static bool matches(const CgvpRowDesc& d) {
        if (d.in_s != IN_S || d.in_v != IN_V || d.onehot != ONEHOT || (d.has_residual_in != 0) != RES_IN ||
            (d.pre_norm != 0) != PRE_NORM || (d.post_residual != 0) != POST_RES || (d.post_norm != 0) != POST_NORM || d.n_gvp != NG)
            return false;
        if (!G0::matches(d.gvp[0])) return false;
        if constexpr (NG == 2) { if (!G1::matches(d.gvp[1])) return false; }
        return true;
    }
};

__device__ __forceinline__ void copy_f4r(float* dst, const float* __restrict__ src, int nfloats) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = threadIdx.x; i < (nfloats >> 2); i += blockDim.x) d4[i] = __ldg(s4 + i);
}

template <class S>
__device__ __forceinline__ void load_ln_params(const CgvpRowArgs& a, float* lnp) {
    if constexpr (S::PRE_NORM)
        for (int i = threadIdx.x; i < S::S0; i += blockDim.x) { lnp[i] = a.ln0_w[i]; lnp[S::S0 + i] = a.ln0_b[i]; }
    if constexpr (S::POST_NORM)
        for (int i = threadIdx.x; i < S::OS; i += blockDim.x) {
            lnp[pad4(2 * S::S0) + i] = a.ln1_w[i];
            lnp[pad4(2 * S::S0) + S::OS + i] = a.ln1_b[i];
        }
}

// Everything the backward pass needs from the forward pass of one row.
template <class S>
struct RowFwd {
    float xs[1][S::S0], xv[3][S::V01];        // program input after the residual add (pre-LayerNorm0)
    float ys[1][S::S0], yv[3][S::V01];        // LayerNorm0 output = chain input
    float cs[1][S::G0::SO], cv[3][S::G0::VO1];   // output of GVP 0 (input of GVP 1)
    float fs[1][S::OS], fv[3][S::OV1];        // pre-LayerNorm1 tensor
    LnStat st0, st1;
};

// READ: the GVPs' s' come from the training stash (a.stash, written by an earlier forward over the same rows); otherwise they
// are computed and, if a.stash is set and `store`, left there.
template <class S, bool READ = false>
__device__ __forceinline__ void rows_forward(const CgvpRowArgs& a, const float* wsm, int w1off, const float* lnp, long long row,
                                             long long rin, RowFwd<S>& f, float (&os)[1][S::OS], float (&ov)[3][S::OV1],
                                             bool store = false) {
    using G0 = typename S::G0;
    float* srow = a.stash ? a.stash + row * S::STASHF : nullptr;
    if constexpr (S::ONEHOT > 0) {
        const int ty = (int)__ldg(a.types + rin);
#pragma unroll
        for (int c = 0; c < S::ONEHOT; ++c) f.xs[0][c] = ty == c ? 1.f : 0.f;
    }
    load_s<S::IN_S, S::ONEHOT>(a.in_s, rin, f.xs);
    load_v<S::IN_V, 0>(a.in_v, rin, f.xv);
    if constexpr (S::RES_IN) {                                             // x + D0(dh), gvp_layers.py:407
        float hs[1][S::IN_S], hv[3][S::V01];
        load_s<S::IN_S, 0>(a.h_s, row, hs);
        load_v<S::IN_V, 0>(a.h_v, row, hv);
        if (a.mask0_s) {
            float m[1][S::IN_S];
            load_s<S::IN_S, 0>(a.mask0_s, row, m);
#pragma unroll
            for (int c = 0; c < S::IN_S; ++c) hs[0][c] *= m[0][c];
        }
        if constexpr (S::IN_V > 0) {
            if (a.mask0_v) {
                float m[1][S::V01];
                load_s<S::IN_V, 0>(a.mask0_v, row, m);
#pragma unroll
                for (int c = 0; c < S::IN_V; ++c)
#pragma unroll
                    for (int p = 0; p < 3; ++p) hv[p][c] *= m[0][c];
            }
        }
#pragma unroll
        for (int c = 0; c < S::IN_S; ++c) f.xs[0][c] += hs[0][c];
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int c = 0; c < S::IN_V; ++c) f.xv[p][c] += hv[p][c];
    }
    if constexpr (S::PRE_NORM) {
        f.st0 = ln_fwd<S::S0, S::V0>(f.xs, f.xv, lnp, lnp + S::S0, f.ys, f.yv);
    } else {
#pragma unroll
        for (int c = 0; c < S::S0; ++c) f.ys[0][c] = f.xs[0][c];
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int c = 0; c < S::V0; ++c) f.yv[p][c] = f.xv[p][c];
    }
    float ls[1][S::OS], lv[3][S::OV1];                                    // chain output
    if constexpr (S::NG == 1) {
        Save<G0> sv;
        if constexpr (READ) { load_s<G0::SO, 0>(srow, 0, sv.sp); gvp_fwd<G0, 1>(wsm, f.ys, f.yv, ls, lv, sv); }
        else { gvp_fwd<G0>(wsm, f.ys, f.yv, ls, lv, sv); if (store && srow) store_s<G0::SO, 0>(srow, 0, sv.sp, false); }
    } else {
        using G1 = typename S::G1;
        {
            Save<G0> sv;
            if constexpr (READ) { load_s<G0::SO, 0>(srow, 0, sv.sp); gvp_fwd<G0, 1>(wsm, f.ys, f.yv, f.cs, f.cv, sv); }
            else { gvp_fwd<G0>(wsm, f.ys, f.yv, f.cs, f.cv, sv); if (store && srow) store_s<G0::SO, 0>(srow, 0, sv.sp, false); }
        }
        {
            Save<G1> sv;
            if constexpr (READ) { load_s<G1::SO, 0>(srow + pad4(G0::SO), 0, sv.sp); gvp_fwd<G1, 1>(wsm + w1off, f.cs, f.cv, ls, lv, sv); }
            else { gvp_fwd<G1>(wsm + w1off, f.cs, f.cv, ls, lv, sv); if (store && srow) store_s<G1::SO, 0>(srow + pad4(G0::SO), 0, sv.sp, false); }
        }
    }
    if constexpr (S::POST_RES) {                                          // x1 + D1(ff(x1)), gvp_layers.py:410
        if (a.mask1_s) {
            float m[1][S::OS];
            load_s<S::OS, 0>(a.mask1_s, row, m);
#pragma unroll
            for (int c = 0; c < S::OS; ++c) ls[0][c] *= m[0][c];
        }
        if constexpr (S::OV > 0) {
            if (a.mask1_v) {
                float m[1][S::OV1];
                load_s<S::OV, 0>(a.mask1_v, row, m);
#pragma unroll
                for (int c = 0; c < S::OV; ++c)
#pragma unroll
                    for (int p = 0; p < 3; ++p) lv[p][c] *= m[0][c];
            }
        }
#pragma unroll
        for (int c = 0; c < S::OS; ++c) f.fs[0][c] = f.ys[0][c] + ls[0][c];
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int c = 0; c < S::OV; ++c) f.fv[p][c] = f.yv[p][c] + lv[p][c];
    } else {
#pragma unroll
        for (int c = 0; c < S::OS; ++c) f.fs[0][c] = ls[0][c];
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int c = 0; c < S::OV; ++c) f.fv[p][c] = lv[p][c];
    }
    if constexpr (S::POST_NORM) {
        f.st1 = ln_fwd<S::OS, S::OV>(f.fs, f.fv, lnp + pad4(2 * S::S0), lnp + pad4(2 * S::S0) + S::OS, os, ov);
    } else {
#pragma unroll
        for (int c = 0; c < S::OS; ++c) os[0][c] = f.fs[0][c];
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int c = 0; c < S::OV; ++c) ov[p][c] = f.fv[p][c];
    }
}

template <class S>
__global__ void __launch_bounds__(256, 2) rows_fwd_reg_kernel(const __grid_constant__ RowsRegArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* wsm = reinterpret_cast<float*>(smem);
    float* lnp = wsm + S::WF;
    copy_f4r(wsm, A.wp[0], S::G0::FWD_FLOATS);
    if constexpr (S::NG == 2) copy_f4r(wsm + S::WF1, A.wp[1], S::G1::FWD_FLOATS);
    load_ln_params<S>(A.a, lnp);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int nw = gridDim.x * nwarp;
    for (int t = blockIdx.x * nwarp + warp; t < A.ntiles; t += nw) {
        const long long row = (long long)t * 32 + lane;
        if (row >= A.a.rows) continue;
        const long long rin = A.a.in_index ? (long long)__ldg(A.a.in_index + row) : row;
        RowFwd<S> f;
        float os[1][S::OS], ov[3][S::OV1];
        rows_forward<S>(A.a, wsm, S::WF1, lnp, row, rin, f, os, ov, true);
        store_s<S::OS, 0>(A.a.out_s, row, os, false);
        if constexpr (S::OV > 0) store_v<S::OV, 0>(A.a.out_v, row, ov, false);
    }
}

template <class S, bool READ>
__global__ void __launch_bounds__(S::BW * 32, 1) rows_bwd_reg_kernel(const __grid_constant__ RowsRegArgs A) {
    using G0 = typename S::G0;
    extern __shared__ __align__(16) unsigned char smem[];
    float* wsm = reinterpret_cast<float*>(smem);
    float* lnp = wsm + S::WT;
    float* arena0 = lnp + S::LNP;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* arena = arena0 + warp * S::PER_WARP;
    float4* stg = reinterpret_cast<float4*>(arena + S::PF);
    copy_f4r(wsm, A.wp[0], S::G0::TOTAL_FLOATS);
    if constexpr (S::NG == 2) copy_f4r(wsm + S::WT1, A.wp[1], S::G1::TOTAL_FLOATS);
    load_ln_params<S>(A.a, lnp);
    for (int i = lane; i < S::PF; i += 32) arena[i] = 0.f;
    __syncthreads();
    const int nw = gridDim.x * S::BW;
    const float one[1][1] = {{1.f}};
    for (int t = blockIdx.x * S::BW + warp; t < A.ntiles; t += nw) {
        const long long row_ = (long long)t * 32 + lane;
        const bool valid = row_ < A.a.rows;
        const long long row = valid ? row_ : (long long)t * 32;            // idle lanes replay the first row
        const long long rin = A.a.in_index ? (long long)__ldg(A.a.in_index + row) : row;
        WarpSink sink{stg, arena, lane, valid};
        RowFwd<S> f;
        float gs[1][S::OS], gv[3][S::OV1];                                // gradient wrt the chain output
        {
            float os[1][S::OS], ov[3][S::OV1];
            rows_forward<S, READ>(A.a, wsm, S::WT1, lnp, row, rin, f, os, ov);
        }
        const float* srow = READ ? A.a.stash + row * S::STASHF : nullptr;
        float dfs[1][S::OS], dfv[3][S::OV1];                              // gradient wrt the pre-LayerNorm1 tensor
        {
            float gys[1][S::OS], gyv[3][S::OV1];
            load_s<S::OS, 0>(A.g.d_out_s, row, gys);
            load_v<S::OV, 0>(A.g.d_out_v, row, gyv);
            if constexpr (S::POST_NORM) {
                float xhat[1][S::OS], b[1][2 * S::OS];
                ln_bwd<S::OS, S::OV>(f.fs, f.fv, f.st1, lnp + pad4(2 * S::S0), gys, gyv, dfs, dfv, xhat);
#pragma unroll
                for (int c = 0; c < S::OS; ++c) { b[0][c] = gys[0][c] * xhat[0][c]; b[0][S::OS + c] = gys[0][c]; }
                sink.template add<1, 2 * S::OS, 1>(S::LN1_OFF, one, b);
            } else {
#pragma unroll
                for (int c = 0; c < S::OS; ++c) dfs[0][c] = gys[0][c];
#pragma unroll
                for (int p = 0; p < 3; ++p)
#pragma unroll
                    for (int c = 0; c < S::OV; ++c) dfv[p][c] = gyv[p][c];
            }
        }
#pragma unroll
        for (int c = 0; c < S::OS; ++c) gs[0][c] = dfs[0][c];
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int c = 0; c < S::OV; ++c) gv[p][c] = dfv[p][c];
        if constexpr (S::POST_RES) {
            if (A.a.mask1_s) {
                float m[1][S::OS];
                load_s<S::OS, 0>(A.a.mask1_s, row, m);
#pragma unroll
                for (int c = 0; c < S::OS; ++c) gs[0][c] *= m[0][c];
            }
            if constexpr (S::OV > 0) {
                if (A.a.mask1_v) {
                    float m[1][S::OV1];
                    load_s<S::OV, 0>(A.a.mask1_v, row, m);
#pragma unroll
                    for (int c = 0; c < S::OV; ++c)
#pragma unroll
                        for (int p = 0; p < 3; ++p) gv[p][c] *= m[0][c];
                }
            }
        }
        // chain backward (each GVP recomputed from its stage input)
        float dys[1][S::S0], dyv[3][S::V01];
        if constexpr (S::NG == 2) {
            using G1 = typename S::G1;
            float g1s[1][G1::SI], g1v[3][G1::VI1];
            {
                Save<G1> sv;
                float so[1][G1::SO], vo[3][G1::VO1], dsin[1][G1::KSD], dvin[3][G1::VI1];
                if constexpr (READ) { load_s<G1::SO, 0>(srow + pad4(G0::SO), 0, sv.sp); gvp_fwd<G1, 1>(wsm + S::WT1, f.cs, f.cv, so, vo, sv); }
                else gvp_fwd<G1>(wsm + S::WT1, f.cs, f.cv, so, vo, sv);
                gvp_bwd<G1>(wsm + S::WT1, sv, f.cs, f.cv, gs, gv, sink, S::GO1, dsin, dvin);
#pragma unroll
                for (int c = 0; c < G1::SI; ++c) g1s[0][c] = dsin[0][c];
#pragma unroll
                for (int p = 0; p < 3; ++p)
#pragma unroll
                    for (int c = 0; c < G1::VI; ++c) g1v[p][c] = dvin[p][c];
            }
            Save<G0> sv;
            float so[1][G0::SO], vo[3][G0::VO1], dsin[1][G0::KSD], dvin[3][G0::VI1];
            if constexpr (READ) { load_s<G0::SO, 0>(srow, 0, sv.sp); gvp_fwd<G0, 1>(wsm, f.ys, f.yv, so, vo, sv); }
            else gvp_fwd<G0>(wsm, f.ys, f.yv, so, vo, sv);
            gvp_bwd<G0>(wsm, sv, f.ys, f.yv, g1s, g1v, sink, S::GO0, dsin, dvin);
#pragma unroll
            for (int c = 0; c < S::S0; ++c) dys[0][c] = dsin[0][c];
#pragma unroll
            for (int p = 0; p < 3; ++p)
#pragma unroll
                for (int c = 0; c < S::V0; ++c) dyv[p][c] = dvin[p][c];
        } else {
            Save<G0> sv;
            float so[1][G0::SO], vo[3][G0::VO1], dsin[1][G0::KSD], dvin[3][G0::VI1];
            if constexpr (READ) { load_s<G0::SO, 0>(srow, 0, sv.sp); gvp_fwd<G0, 1>(wsm, f.ys, f.yv, so, vo, sv); }
            else gvp_fwd<G0>(wsm, f.ys, f.yv, so, vo, sv);
            // an embed of leaf features (no residual, no pre-norm, nobody asks for d_in) needs weight gradients only
            if (S::PRE_NORM || S::POST_RES || S::RES_IN || A.g.d_in_s || A.g.d_in_v)
                gvp_bwd<G0>(wsm, sv, f.ys, f.yv, gs, gv, sink, S::GO0, dsin, dvin);
            else
                gvp_bwd<G0, decltype(sink), false>(wsm, sv, f.ys, f.yv, gs, gv, sink, S::GO0, dsin, dvin);
#pragma unroll
            for (int c = 0; c < S::S0; ++c) dys[0][c] = dsin[0][c];
#pragma unroll
            for (int p = 0; p < 3; ++p)
#pragma unroll
                for (int c = 0; c < S::V0; ++c) dyv[p][c] = S::V0 > 0 ? dvin[p][c] : 0.f;
        }
        if constexpr (S::POST_RES) {                                      // residual branch
#pragma unroll
            for (int c = 0; c < S::S0; ++c) dys[0][c] += dfs[0][c];
#pragma unroll
            for (int p = 0; p < 3; ++p)
#pragma unroll
                for (int c = 0; c < S::V0; ++c) dyv[p][c] += dfv[p][c];
        }
        float dxs[1][S::S0], dxv[3][S::V01];
        if constexpr (S::PRE_NORM) {
            float xhat[1][S::S0], b[1][2 * S::S0];
            ln_bwd<S::S0, S::V0>(f.xs, f.xv, f.st0, lnp, dys, dyv, dxs, dxv, xhat);
#pragma unroll
            for (int c = 0; c < S::S0; ++c) { b[0][c] = dys[0][c] * xhat[0][c]; b[0][S::S0 + c] = dys[0][c]; }
            sink.template add<1, 2 * S::S0, 1>(S::LN0_OFF, one, b);
        } else {
#pragma unroll
            for (int c = 0; c < S::S0; ++c) dxs[0][c] = dys[0][c];
#pragma unroll
            for (int p = 0; p < 3; ++p)
#pragma unroll
                for (int c = 0; c < S::V0; ++c) dxv[p][c] = dyv[p][c];
        }
        if (valid) {
            if (A.g.d_in_s) store_s<S::IN_S, S::ONEHOT>(A.g.d_in_s, rin, dxs, false);
            if constexpr (S::IN_V > 0) { if (A.g.d_in_v) store_v<S::IN_V, 0>(A.g.d_in_v, rin, dxv, false); }
            if constexpr (S::RES_IN) {
                if (A.g.d_h_s) {
                    float hs[1][S::IN_S], hv[3][S::V01];
#pragma unroll
                    for (int c = 0; c < S::IN_S; ++c) hs[0][c] = dxs[0][c];
#pragma unroll
                    for (int p = 0; p < 3; ++p)
#pragma unroll
                        for (int c = 0; c < S::IN_V; ++c) hv[p][c] = dxv[p][c];
                    if (A.a.mask0_s) {
                        float m[1][S::IN_S];
                        load_s<S::IN_S, 0>(A.a.mask0_s, row, m);
#pragma unroll
                        for (int c = 0; c < S::IN_S; ++c) hs[0][c] *= m[0][c];
                    }
                    if constexpr (S::IN_V > 0) {
                        if (A.a.mask0_v) {
                            float m[1][S::V01];
                            load_s<S::IN_V, 0>(A.a.mask0_v, row, m);
#pragma unroll
                            for (int c = 0; c < S::IN_V; ++c)
#pragma unroll
                                for (int p = 0; p < 3; ++p) hv[p][c] *= m[0][c];
                        }
                    }
                    store_s<S::IN_S, 0>(A.g.d_h_s, row, hs, false);
                    if constexpr (S::IN_V > 0) store_v<S::IN_V, 0>(A.g.d_h_v, row, hv, false);
                }
            }
        }
    }
    __syncthreads();
    float* out = A.partial + (long long)blockIdx.x * S::PF;
    for (int i = threadIdx.x; i < S::PF; i += blockDim.x) {
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < S::BW; ++w) sum += arena0[w * S::PER_WARP + i];
        out[i] = sum;
    }
}

bool cgvp_fast_paths_enabled();
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <class S>
static bool buffers_ok(const CgvpRowArgs* a, const CgvpRowGradArgs* g) {
    bool ok = aligned16(a->in_s) && aligned16(a->in_v) && aligned16(a->h_s) && aligned16(a->h_v) && aligned16(a->mask0_s) &&
              aligned16(a->mask0_v) && aligned16(a->mask1_s) && aligned16(a->mask1_v) && aligned16(a->out_s) && aligned16(a->out_v);
    if (g) ok = ok && aligned16(g->d_out_s) && aligned16(g->d_out_v) && aligned16(g->d_in_s) && aligned16(g->d_in_v) &&
                aligned16(g->d_h_s) && aligned16(g->d_h_v);
    return ok;
}

template <class S>
static int launch_fwd(const CgvpRowArgs* args, cudaStream_t st) {
    RowsRegArgs A;
    memset(&A, 0, sizeof(A));
    A.a = *args;
    A.wp[0] = args->h_packed[0];
    if (S::NG == 2) A.wp[1] = args->h_packed[1];
    A.ntiles = (int)cdiv64(args->rows, 32);
    if (A.a.stash && !aligned16(A.a.stash)) A.a.stash = nullptr;
    const int sms = cgvp_num_sms();
    const int grid = (int)min((long long)cdiv(A.ntiles, 8), (long long)sms * 2);
    const size_t smem = S::smem_fwd();
    CGVP_CUDA(cudaFuncSetAttribute(rows_fwd_reg_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cgvp_prof_begin(CGVP_K_ROWS_FWD, st);
    rows_fwd_reg_kernel<S><<<grid, 256, smem, st>>>(A);
    cgvp_prof_end(CGVP_K_ROWS_FWD, st);
    CGVP_LAUNCH_CHECK("rows_fwd_reg_kernel");
    return 0;
}

template <class S>
static int launch_bwd(const CgvpRowArgs* args, const CgvpRowGradArgs* grads, void* ws, int64_t ws_bytes, cudaStream_t st) {
    RowsRegArgs A;
    memset(&A, 0, sizeof(A));
    A.a = *args;
    A.g = *grads;
    A.wp[0] = args->h_packed[0];
    if (S::NG == 2) A.wp[1] = args->h_packed[1];
    A.ntiles = (int)cdiv64(args->rows, 32);
    const int sms = cgvp_num_sms();
    int grid = (int)min((long long)cdiv(A.ntiles, S::BW), (long long)sms);
    if (grid < 1) grid = 1;
    const int64_t need = (int64_t)(grid + 1) * S::PF * 4;
    CGVP_REQUIRE(ws && ws_bytes >= need && aligned16(ws), "rows_bwd: workspace too small (%lld < %lld)", (long long)ws_bytes,
                 (long long)need);
    A.partial = reinterpret_cast<float*>(ws);
    float* reduced = A.partial + (int64_t)grid * S::PF;
    const size_t smem = S::smem_bwd();
    const bool read = args->stash != nullptr && aligned16(args->stash);
    if (!read) A.a.stash = nullptr;
    if (read) CGVP_CUDA(cudaFuncSetAttribute(rows_bwd_reg_kernel<S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else CGVP_CUDA(cudaFuncSetAttribute(rows_bwd_reg_kernel<S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cgvp_prof_begin(CGVP_K_ROWS_BWD, st);
    if (read) rows_bwd_reg_kernel<S, true><<<grid, S::BW * 32, smem, st>>>(A);
    else rows_bwd_reg_kernel<S, false><<<grid, S::BW * 32, smem, st>>>(A);
    cgvp_prof_end(CGVP_K_ROWS_BWD, st);
    CGVP_LAUNCH_CHECK("rows_bwd_reg_kernel");
    CgvpSeg seg[CGVP_MAX_SEGS];
    memset(seg, 0, sizeof(seg));
    int ns = 0;
    seg[ns].dst = grads->h_packed_grads[0]; seg[ns].off = S::GO0; seg[ns].n = S::G0::FWD_FLOATS; ++ns;
    if (S::NG == 2) { seg[ns].dst = grads->h_packed_grads[1]; seg[ns].off = S::GO1; seg[ns].n = S::G1::FWD_FLOATS; ++ns; }
    if (S::PRE_NORM) {
        seg[ns].dst = grads->d_ln0_w; seg[ns].off = S::LN0_OFF; seg[ns].n = S::S0; ++ns;
        seg[ns].dst = grads->d_ln0_b; seg[ns].off = S::LN0_OFF + S::S0; seg[ns].n = S::S0; ++ns;
    }
    if (S::POST_NORM) {
        seg[ns].dst = grads->d_ln1_w; seg[ns].off = S::LN1_OFF; seg[ns].n = S::OS; ++ns;
        seg[ns].dst = grads->d_ln1_b; seg[ns].off = S::LN1_OFF + S::OS; seg[ns].n = S::OS; ++ns;
    }
    return cgvp_reduce_partials(A.partial, grid, S::PF, reduced, seg, ns, st);
}

template <class S>
static bool try_fwd(const CgvpRowDesc* d, const CgvpRowArgs* a, cudaStream_t st, int* rc) {
    if (!S::matches(*d) || !buffers_ok<S>(a, nullptr)) return false;
    *rc = launch_fwd<S>(a, st);
    return true;
}
template <class S>
static bool try_bwd(const CgvpRowDesc* d, const CgvpRowArgs* a, const CgvpRowGradArgs* g, void* ws, int64_t wsb, cudaStream_t st, int* rc) {
    if (!S::matches(*d) || !buffers_ok<S>(a, g)) return false;
    *rc = launch_bwd<S>(a, g, ws, wsb, st);
    return true;
}


// One translation unit per instance (rows_reg_<name>.cu) keeps the build parallel: each exports these three functions.
#define CGVP_ROWS_INSTANCE(NAME, SPEC)                                                                                       \
    bool rows_try_fwd_##NAME(const CgvpRowDesc* d, const CgvpRowArgs* a, cudaStream_t st, int* rc) { return try_fwd<SPEC>(d, a, st, rc); } \
    bool rows_try_bwd_##NAME(const CgvpRowDesc* d, const CgvpRowArgs* a, const CgvpRowGradArgs* g, void* ws, int64_t wsb,   \
                             cudaStream_t st, int* rc) { return try_bwd<SPEC>(d, a, g, ws, wsb, st, rc); }                   \
    int rows_pf_##NAME(const CgvpRowDesc* d) { return SPEC::matches(*d) ? SPEC::PF : 0; }                                    \
    int rows_stashf_##NAME(const CgvpRowDesc* d) { return SPEC::matches(*d) ? SPEC::STASHF : 0; }

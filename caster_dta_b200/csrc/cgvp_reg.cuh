// Register-resident, compile-time specialised GVP arithmetic (sm_100a).
//
// One thread owns one row (edge or node) and keeps the WHOLE row state in registers: every loop below has
// compile-time bounds and is fully unrolled, so the arrays never touch local memory; weights are read from the
// packed block (cgvp_common.cuh) at warp-uniform addresses (LDS.128 broadcast) and fed to packed fp32x2 FMAs.
// Compared with the generic shared-memory tile kernels (cgvp_tile.cuh) the per-row shared-memory footprint drops
// from ~1.5 KB to zero, which lifts occupancy from 4 to 8-16 warps per SM and removes the CTA-wide phase barriers.
//
// The maths restates models/gvp_layers.py:142-175 (GVP), :231-242 (LayerNorm) and SURVEY.md Appendix E (backward).
// Everything is __host__ __device__ so that tests/ can run the very same templates on the CPU against the oracle.
#pragma once
#include "cgvp_common.cuh"

namespace cgvpr {

constexpr CGVP_HD inline int pad4(int x) { return (x + 3) / 4 * 4; }
constexpr CGVP_HD inline int max1(int x) { return x > 0 ? x : 1; }

// ---- compile-time description of one GVP (mirrors CgvpGvpDesc / GvpP) -------------------------------------------
template <int SI_, int VI_, int SO_, int VO_, int H_, int SACT_, int VACT_, int GATE_>
struct GvpC {
    static constexpr int SI = SI_, VI = VI_, SO = SO_, VO = VO_, H = (VI_ > 0 ? H_ : 0);
    static constexpr int SACT = SACT_, VACT = VACT_;
    static constexpr bool GATE = GATE_ != 0 && VI_ > 0 && VO_ > 0;
    static constexpr int VI1 = max1(VI), VO1 = max1(VO), H1 = max1(H);
    static constexpr int VIP = pad4(VI), HP = pad4(H), SOP = pad4(SO), VOP = pad4(VO);
    static constexpr int KSD = SI + H;                 // data width of the ws input [s ; vn]
    static constexpr int KS = KSD + 1;                 // + ones column (bias row)
    static constexpr int KSDP = pad4(KSD);
    static constexpr int KSV = SO + 1;                 // gate input [act(s') ; 1]
    static constexpr CgvpGvpDesc desc() { return CgvpGvpDesc{SI_, VI_, SO_, VO_, H_, SACT_, VACT_, GATE_}; }
    static constexpr int O_WH_T = make_gvp_p(desc()).o_wh_t;
    static constexpr int O_WS_T = make_gvp_p(desc()).o_ws_t;
    static constexpr int O_WV_T = make_gvp_p(desc()).o_wv_t;
    static constexpr int O_WSV_T = make_gvp_p(desc()).o_wsv_t;
    static constexpr int O_WH_B = make_gvp_p(desc()).o_wh_b;
    static constexpr int O_WS_B = make_gvp_p(desc()).o_ws_b;
    static constexpr int O_WV_B = make_gvp_p(desc()).o_wv_b;
    static constexpr int O_WSV_B = make_gvp_p(desc()).o_wsv_b;
    static constexpr int FWD_FLOATS = make_gvp_p(desc()).fwd_floats;
    static constexpr int TOTAL_FLOATS = make_gvp_p(desc()).total_floats;
    static bool matches(const CgvpGvpDesc& d) {
        const GvpP a = make_gvp_p(d), b = make_gvp_p(desc());
        return a.si == b.si && a.vi == b.vi && a.so == b.so && a.vo == b.vo && a.h == b.h && a.sact == b.sact &&
               a.vact == b.vact && a.gate == b.gate;
    }
};

// ---- scalar helpers ---------------------------------------------------------------------------------------------
template <int ACT>
CGVP_HD inline float actf(float x) {
    if constexpr (ACT == CGVP_ACT_RELU) return fmaxf(x, 0.f);
    else if constexpr (ACT == CGVP_ACT_SIGMOID) return 1.f / (1.f + expf(-x));
    else return x;
}
// derivative expressed through the activation OUTPUT y
template <int ACT>
CGVP_HD inline float actb(float y) {
    if constexpr (ACT == CGVP_ACT_RELU) return y > 0.f ? 1.f : 0.f;
    else if constexpr (ACT == CGVP_ACT_SIGMOID) return y * (1.f - y);
    else return 1.f;
}
CGVP_HD inline float sigm(float x) { return 1.f / (1.f + expf(-x)); }

// c += a * b on two lanes (FFMA2 on sm_100a)
CGVP_HD inline void fma2(float2& c, float a, float2 b) {
#ifdef __CUDA_ARCH__
    unsigned long long cc, aa, bb;
    asm("mov.b64 %0, {%1, %2};" : "=l"(cc) : "f"(c.x), "f"(c.y));
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b.x), "f"(b.y));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(cc) : "l"(aa), "l"(bb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(c.x), "=f"(c.y) : "l"(cc));
#else
    c.x = fmaf(a, b.x, c.x);
    c.y = fmaf(a, b.y, c.y);
#endif
}

CGVP_HD inline float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

template <int P, int N>
CGVP_HD inline void zero2(float (&a)[P][N]) {
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
        for (int i = 0; i < N; ++i) a[p][i] = 0.f;
}

// ---- the row-linear primitive -----------------------------------------------------------------------------------
// y[p][Y0 + o] += sum_{k<K} x[p][X0 + k] * W[k * LD + o]     o < N, p < PL  (rows of W are zero padded to LD)
// Outputs are produced in register blocks of OB4 float4 columns: PL * OB4 * 2 independent FFMA2 chains.
template <int K, int N, int LD, int X0, int Y0, int PL, int XN, int YN>
CGVP_HD inline void mv(const float* __restrict__ W, const float (&x)[PL][XN], float (&y)[PL][YN]) {
    static_assert(X0 + K <= XN && Y0 + N <= YN, "mv: operand slice out of range");
    constexpr int N4 = (N + 3) / 4;
    constexpr int OB4 = PL == 1 ? 4 : 2;
#pragma unroll
    for (int b0 = 0; b0 < N4; b0 += OB4) {
        float2 acc[PL][OB4 * 2];
#pragma unroll
        for (int p = 0; p < PL; ++p)
#pragma unroll
            for (int j = 0; j < OB4 * 2; ++j) {
                const int o = 4 * b0 + 2 * j;
                acc[p][j].x = o < N ? y[p][Y0 + (o < N ? o : 0)] : 0.f;
                acc[p][j].y = o + 1 < N ? y[p][Y0 + (o + 1 < N ? o + 1 : 0)] : 0.f;
            }
#pragma unroll
        for (int k = 0; k < K; ++k) {
#pragma unroll
            for (int j = 0; j < OB4; ++j) {
                if (b0 + j < N4) {
                    const float4 w = ld4(W + k * LD + 4 * (b0 + j));
#pragma unroll
                    for (int p = 0; p < PL; ++p) {
                        fma2(acc[p][2 * j], x[p][X0 + k], make_float2(w.x, w.y));
                        fma2(acc[p][2 * j + 1], x[p][X0 + k], make_float2(w.z, w.w));
                    }
                }
            }
        }
#pragma unroll
        for (int p = 0; p < PL; ++p)
#pragma unroll
            for (int j = 0; j < OB4 * 2; ++j) {
                const int o = 4 * b0 + 2 * j;
                if (o < N) y[p][Y0 + (o < N ? o : 0)] = acc[p][j].x;
                if (o + 1 < N) y[p][Y0 + (o + 1 < N ? o + 1 : 0)] = acc[p][j].y;
            }
    }
}

// ---- GVP forward (gvp_layers.py:142-175) ------------------------------------------------------------------------
template <class G>
struct Save {
    float vh[3][G::H1];     // hidden vectors                                                  :152
    float vn[1][G::H1];     // their clamped norms                                             :153
    float sp[1][G::SO];     // s' BEFORE the scalar activation                                 :154
    float vo[3][G::VO1];    // output vectors BEFORE gating / vector activation               :156
    float sg[1][G::VO1];    // the factor they were multiplied with                            :163 / :166
};

// SPM (how s' = W_s [s ; vn] + b is formed):
//   0  computed here from all of `s`;
//   1  sv.sp already holds s' (e.g. read back from a training stash): the W_s projection -- most of a GVP's scalar
//      arithmetic -- is skipped and `s` is not read; everything else (Vh, norms, Vo, gate) is recomputed from v and s';
//   2  sv.sp already holds the bias and the contribution of the scalar inputs OUTSIDE [S0, S0 + SN) (e.g. the per-node
//      projections of the source / target scalars of a message, computed once per node instead of once per edge): only
//      rows [S0, S0 + SN) of `s` and the norm rows are added here.
template <class G, int SPM = 0, int S0 = 0, int SN = G::SI>
CGVP_HD inline void gvp_fwd(const float* __restrict__ W, const float (&s)[1][G::SI], const float (&v)[3][G::VI1],
                            float (&so)[1][G::SO], float (&vout)[3][G::VO1], Save<G>& sv) {
    if constexpr (G::VI > 0) {
        zero2(sv.vh);
        mv<G::VI, G::H, G::HP, 0, 0>(W + G::O_WH_T, v, sv.vh);
#pragma unroll
        for (int o = 0; o < G::H; ++o) {
            const float q = sv.vh[0][o] * sv.vh[0][o] + sv.vh[1][o] * sv.vh[1][o] + sv.vh[2][o] * sv.vh[2][o];
            sv.vn[0][o] = sqrtf(fmaxf(q, CGVP_EPS));
        }
    }
    if constexpr (SPM == 0) {
#pragma unroll
        for (int o = 0; o < G::SO; ++o) sv.sp[0][o] = W[G::O_WS_T + G::KSD * G::SOP + o];      // bias row
        mv<G::SI, G::SO, G::SOP, 0, 0>(W + G::O_WS_T, s, sv.sp);
        if constexpr (G::VI > 0) mv<G::H, G::SO, G::SOP, 0, 0>(W + G::O_WS_T + G::SI * G::SOP, sv.vn, sv.sp);
    } else if constexpr (SPM == 2) {
        static_assert(S0 >= 0 && S0 + SN <= G::SI, "gvp_fwd: scalar slice out of range");
        mv<SN, G::SO, G::SOP, S0, 0>(W + G::O_WS_T + S0 * G::SOP, s, sv.sp);
        if constexpr (G::VI > 0) mv<G::H, G::SO, G::SOP, 0, 0>(W + G::O_WS_T + G::SI * G::SOP, sv.vn, sv.sp);
    }
#pragma unroll
    for (int o = 0; o < G::SO; ++o) so[0][o] = actf<G::SACT>(sv.sp[0][o]);                  // :172-173
    if constexpr (G::VO > 0) {
        if constexpr (G::VI > 0) {
            zero2(sv.vo);
            mv<G::H, G::VO, G::VOP, 0, 0>(W + G::O_WV_T, sv.vh, sv.vo);                     // :156
            if constexpr (G::GATE) {                                                        // :158-163
                float gi[1][G::SO], g[1][G::VO];
#pragma unroll
                for (int o = 0; o < G::SO; ++o) gi[0][o] = actf<G::VACT>(sv.sp[0][o]);
#pragma unroll
                for (int o = 0; o < G::VO; ++o) g[0][o] = W[G::O_WSV_T + G::SO * G::VOP + o];
                mv<G::SO, G::VO, G::VOP, 0, 0>(W + G::O_WSV_T, gi, g);
#pragma unroll
                for (int o = 0; o < G::VO; ++o) sv.sg[0][o] = sigm(g[0][o]);
            } else {
#pragma unroll
                for (int o = 0; o < G::VO; ++o) {
                    if constexpr (G::VACT != CGVP_ACT_NONE) {                               // :164-166
                        const float q = sv.vo[0][o] * sv.vo[0][o] + sv.vo[1][o] * sv.vo[1][o] + sv.vo[2][o] * sv.vo[2][o];
                        sv.sg[0][o] = actf<G::VACT>(sqrtf(fmaxf(q, CGVP_EPS)));
                    } else {
                        sv.sg[0][o] = 1.f;
                    }
                }
            }
#pragma unroll
            for (int p = 0; p < 3; ++p)
#pragma unroll
                for (int o = 0; o < G::VO; ++o) vout[p][o] = sv.vo[p][o] * sv.sg[0][o];
        } else {
            zero2(vout);                                                                    // :169-171
        }
    }
}

// ---- GVP backward ------------------------------------------------------------------------
// `sink.add<KA, NB, NP>(offset, A, B)` receives the operand pair of each weight-gradient GEMM,
//     G[offset + a * pad4(NB) + b] += sum_p A[p][a] * B[p][b]          (layout of the forward packed block),
// interleaved with the data path so that operands die as early as possible.
// On entry gs / gv hold the gradient of the GVP outputs; on exit dsin = [dS_in ; dvn] and dvin = dV_in.
// NEED_DX = false (the GVP reads leaf data, e.g. the raw edge / node features): only the weight gradients are wanted,
// so dS_in and dV_in are not formed (dvn still is: it feeds dW_h).
// Slice mode (S0, SN with SN < SI; the partner of gvp_fwd<G, 2, S0, SN>): only rows [S0, S0 + SN) of the scalar input are
// handled per row -- their weight-gradient rows, their dS_in columns -- plus the norm rows; the remaining scalar inputs and
// the bias are handled by the caller from `ds_out` = dL/ds' (e.g. reduced per node first: W^T (sum_e ds'_e), (sum_e ds'_e) x s_n).
template <class G, class Sink, bool NEED_DX = true, int S0 = 0, int SN = G::SI>
CGVP_HD inline void gvp_bwd_ds(const float* __restrict__ W, const Save<G>& sv, const float (&s)[1][G::SI],
                               const float (&v)[3][G::VI1], const float (&gs)[1][G::SO], const float (&gv)[3][G::VO1],
                               Sink& sink, int goff, float (&dsin)[1][G::KSD], float (&dvin)[3][G::VI1],
                               float (&ds)[1][G::SO]) {
    constexpr bool HASV = G::VI > 0 && G::VO > 0;
    constexpr bool SLICE = SN < G::SI;
    static_assert(S0 >= 0 && S0 + SN <= G::SI && (!SLICE || S0 % 4 == 0), "gvp_bwd: scalar slice");
    float dvo[3][G::VO1], dg[1][G::VO1];
    if constexpr (HASV) {
#pragma unroll
        for (int o = 0; o < G::VO; ++o) {
            const float sg = sv.sg[0][o];
            const float dot = gv[0][o] * sv.vo[0][o] + gv[1][o] * sv.vo[1][o] + gv[2][o] * sv.vo[2][o];
            if constexpr (G::GATE) {
                dg[0][o] = dot * sg * (1.f - sg);
#pragma unroll
                for (int p = 0; p < 3; ++p) dvo[p][o] = gv[p][o] * sg;
            } else if constexpr (G::VACT != CGVP_ACT_NONE) {
                const float q = sv.vo[0][o] * sv.vo[0][o] + sv.vo[1][o] * sv.vo[1][o] + sv.vo[2][o] * sv.vo[2][o];
                const float t = q >= CGVP_EPS ? dot * actb<G::VACT>(sg) / sqrtf(q) : 0.f;
#pragma unroll
                for (int p = 0; p < 3; ++p) dvo[p][o] = gv[p][o] * sg + sv.vo[p][o] * t;
            } else {
#pragma unroll
                for (int p = 0; p < 3; ++p) dvo[p][o] = gv[p][o];
            }
        }
    }
    // ds' = dS_out * sact'(s_out) + (dg . wsv) * vact'(gate input)
#pragma unroll
    for (int o = 0; o < G::SO; ++o) ds[0][o] = gs[0][o] * actb<G::SACT>(actf<G::SACT>(sv.sp[0][o]));
    if constexpr (G::GATE) {
        float t[1][G::SO], gi1[1][G::KSV];
        zero2(t);
        mv<G::VO, G::SO, G::SOP, 0, 0>(W + G::O_WSV_B, dg, t);
#pragma unroll
        for (int o = 0; o < G::SO; ++o) {
            const float gi = actf<G::VACT>(sv.sp[0][o]);
            gi1[0][o] = gi;
            ds[0][o] += t[0][o] * actb<G::VACT>(gi);
        }
        gi1[0][G::SO] = 1.f;
        sink.template add<G::KSV, G::VO, 1>(goff + G::O_WSV_T, gi1, dg);
    }
    if constexpr (!SLICE) {
        float a[1][G::KS];
#pragma unroll
        for (int k = 0; k < G::SI; ++k) a[0][k] = s[0][k];
        if constexpr (G::VI > 0) {
#pragma unroll
            for (int k = 0; k < G::H; ++k) a[0][G::SI + k] = sv.vn[0][k];
        }
        a[0][G::KSD] = 1.f;
        sink.template add<G::KS, G::SO, 1>(goff + G::O_WS_T, a, ds);
    } else {
        float a[1][SN];
#pragma unroll
        for (int k = 0; k < SN; ++k) a[0][k] = s[0][S0 + k];
        sink.template add<SN, G::SO, 1>(goff + G::O_WS_T + S0 * G::SOP, a, ds);
        if constexpr (G::VI > 0) sink.template add<G::H, G::SO, 1>(goff + G::O_WS_T + G::SI * G::SOP, sv.vn, ds);
    }
    zero2(dsin);
    if constexpr (SLICE) {
        constexpr int C0 = G::SI / 4 * 4;                                   // dvn columns (from a 16-byte boundary)
        static_assert(S0 + SN <= C0, "gvp_bwd: scalar slice overlaps the norm columns' block");
        if constexpr (NEED_DX) mv<G::SO, SN, G::KSDP, 0, S0>(W + G::O_WS_B + S0, ds, dsin);
        if constexpr (G::VI > 0) mv<G::SO, G::KSD - C0, G::KSDP, 0, C0>(W + G::O_WS_B + C0, ds, dsin);
    } else if constexpr (NEED_DX || G::VI == 0) {
        if constexpr (NEED_DX) mv<G::SO, G::KSD, G::KSDP, 0, 0>(W + G::O_WS_B, ds, dsin);   // [dS_in ; dvn] = ds' . ws
    } else {
        constexpr int C0 = G::SI / 4 * 4;                                   // dvn columns only (from a 16-byte boundary)
        mv<G::SO, G::KSD - C0, G::KSDP, 0, C0>(W + G::O_WS_B + C0, ds, dsin);
    }
    if constexpr (G::VI > 0) {
        float dvh[3][G::H1];
        zero2(dvh);
        if constexpr (G::VO > 0) {
            mv<G::VO, G::H, G::HP, 0, 0>(W + G::O_WV_B, dvo, dvh);          // wv^T dVo
            sink.template add<G::H, G::VO, 3>(goff + G::O_WV_T, sv.vh, dvo);
        }
#pragma unroll
        for (int k = 0; k < G::H; ++k) {                                    // + Vh * dvn / vn where the clamp passes
            const float q = sv.vh[0][k] * sv.vh[0][k] + sv.vh[1][k] * sv.vh[1][k] + sv.vh[2][k] * sv.vh[2][k];
            const float f = q >= CGVP_EPS ? dsin[0][G::SI + k] / sv.vn[0][k] : 0.f;
#pragma unroll
            for (int p = 0; p < 3; ++p) dvh[p][k] += sv.vh[p][k] * f;
        }
        sink.template add<G::VI, G::H, 3>(goff + G::O_WH_T, v, dvh);
        zero2(dvin);
        if constexpr (NEED_DX) mv<G::H, G::VI, G::VIP, 0, 0>(W + G::O_WH_B, dvh, dvin);     // dV_in = wh^T dVh
    }
}

template <class G, class Sink, bool NEED_DX = true>
CGVP_HD inline void gvp_bwd(const float* __restrict__ W, const Save<G>& sv, const float (&s)[1][G::SI],
                            const float (&v)[3][G::VI1], const float (&gs)[1][G::SO], const float (&gv)[3][G::VO1],
                            Sink& sink, int goff, float (&dsin)[1][G::KSD], float (&dvin)[3][G::VI1]) {
    float ds[1][G::SO];
    gvp_bwd_ds<G, Sink, NEED_DX>(W, sv, s, v, gs, gv, sink, goff, dsin, dvin, ds);
}

// ---- LayerNorm (gvp_layers.py:231-242) --------------------------------------------------------------------------
struct LnStat { float mean, rstd, rms; };

template <int S, int C>
CGVP_HD inline LnStat ln_fwd(const float (&xs)[1][S], const float (&xv)[3][max1(C)], const float* __restrict__ w,
                             const float* __restrict__ b, float (&ys)[1][S], float (&yv)[3][max1(C)]) {
    LnStat st;
    float mean = 0.f;
#pragma unroll
    for (int k = 0; k < S; ++k) mean += xs[0][k];
    mean /= (float)S;
    float var = 0.f;
#pragma unroll
    for (int k = 0; k < S; ++k) { const float d = xs[0][k] - mean; var += d * d; }
    const float rstd = rsqrtf(var / (float)S + CGVP_LN_EPS);
#pragma unroll
    for (int k = 0; k < S; ++k) ys[0][k] = (xs[0][k] - mean) * rstd * w[k] + b[k];
    st.mean = mean; st.rstd = rstd; st.rms = 1.f;
    if constexpr (C > 0) {
        float m = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) m += fmaxf(xv[0][c] * xv[0][c] + xv[1][c] * xv[1][c] + xv[2][c] * xv[2][c], CGVP_EPS);   // :240
        const float rms = sqrtf(m / (float)C);                                                                             // :241
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int c = 0; c < C; ++c) yv[p][c] = xv[p][c] / rms;
        st.rms = rms;
    }
    return st;
}

// dy -> dx (in place allowed: dxs may alias dys).  xhat[k] = (x - mean) * rstd is returned for the parameter gradients.
template <int S, int C>
CGVP_HD inline void ln_bwd(const float (&xs)[1][S], const float (&xv)[3][max1(C)], const LnStat& st,
                           const float* __restrict__ w, const float (&dys)[1][S], const float (&dyv)[3][max1(C)],
                           float (&dxs)[1][S], float (&dxv)[3][max1(C)], float (&xhat)[1][S]) {
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int k = 0; k < S; ++k) {
        xhat[0][k] = (xs[0][k] - st.mean) * st.rstd;
        const float dyh = dys[0][k] * w[k];
        m1 += dyh; m2 += dyh * xhat[0][k];
    }
    m1 /= (float)S; m2 /= (float)S;
#pragma unroll
    for (int k = 0; k < S; ++k) dxs[0][k] = st.rstd * (dys[0][k] * w[k] - m1 - xhat[0][k] * m2);
    if constexpr (C > 0) {
        float dot = 0.f;
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int c = 0; c < C; ++c) dot += dyv[p][c] * xv[p][c];
        const float coef = dot / ((float)C * st.rms * st.rms * st.rms);
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float q = xv[0][c] * xv[0][c] + xv[1][c] * xv[1][c] + xv[2][c] * xv[2][c];
            const float pass = q >= CGVP_EPS ? coef : 0.f;
#pragma unroll
            for (int p = 0; p < 3; ++p) dxv[p][c] = dyv[p][c] / st.rms - xv[p][c] * pass;
        }
    }
}

// weight-gradient sink that ignores everything (forward-only instantiations never call it)
struct NullSink {
    template <int KA, int NB, int NP, int AX, int BX>
    CGVP_HD inline void add(int, const float (&)[NP][AX], const float (&)[NP][BX]) {}
};

// plain accumulation into one arena (host tests, and the reference semantics of every device sink)
struct DirectSink {
    float* G;
    template <int KA, int NB, int NP, int AX, int BX>
    CGVP_HD inline void add(int off, const float (&A)[NP][AX], const float (&B)[NP][BX]) {
        for (int a = 0; a < KA; ++a)
            for (int b = 0; b < NB; ++b) {
                float t = 0.f;
                for (int p = 0; p < NP; ++p) t += A[p][a] * B[p][b];
                G[off + a * pad4(NB) + b] += t;
            }
    }
};

}  // namespace cgvpr

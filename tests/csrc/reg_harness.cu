// TEST INFRASTRUCTURE: runs the register-resident GVP / LayerNorm templates of caster_dta_b200/csrc/cgvp_reg.cuh on the
// HOST (they are __host__ __device__), so their maths and packed-weight offsets can be checked against the oracle
// without a GPU.  Never linked into libcastergvp.so.
#include "../../caster_dta_b200/csrc/cgvp_reg.cuh"

using namespace cgvpr;

template <class G>
static void run_gvp(int n, const float* W, const float* s, const float* v, const float* gs, const float* gv, float* so,
                    float* vo, float* dsin, float* dvin, float* Gacc) {
    for (int r = 0; r < n; ++r) {
        float xs[1][G::SI], xv[3][G::VI1], ys[1][G::SO], yv[3][G::VO1];
        for (int k = 0; k < G::SI; ++k) xs[0][k] = s[r * G::SI + k];
        for (int c = 0; c < G::VI; ++c)
            for (int p = 0; p < 3; ++p) xv[p][c] = v[(r * G::VI + c) * 3 + p];
        Save<G> sv;
        gvp_fwd<G>(W, xs, xv, ys, yv, sv);
        for (int k = 0; k < G::SO; ++k) so[r * G::SO + k] = ys[0][k];
        for (int c = 0; c < G::VO; ++c)
            for (int p = 0; p < 3; ++p) vo[(r * G::VO + c) * 3 + p] = yv[p][c];
        float g1[1][G::SO], g3[3][G::VO1], di[1][G::KSD], dv[3][G::VI1];
        for (int k = 0; k < G::SO; ++k) g1[0][k] = gs[r * G::SO + k];
        for (int c = 0; c < G::VO; ++c)
            for (int p = 0; p < 3; ++p) g3[p][c] = gv[(r * G::VO + c) * 3 + p];
        DirectSink sink{Gacc};
        gvp_bwd<G>(W, sv, xs, xv, g1, g3, sink, 0, di, dv);
        for (int k = 0; k < G::SI; ++k) dsin[r * G::SI + k] = di[0][k];
        for (int c = 0; c < G::VI; ++c)
            for (int p = 0; p < 3; ++p) dvin[(r * G::VI + c) * 3 + p] = dv[p][c];
    }
}

// The split path of message GVP 0 (conv_reg.cu, round 2): the scalar input is [s_j (NS) ; e_s (ES) ; s_i (NS)]; the node
// slices go through per-node projections (forward: s' starts from W_sj s_j + W_si s_i + b) and, in the backward, through
// ds' (d s_j = W_sj^T ds', dW_s[node rows] += s (x) ds', bias += ds'); the template handles the edge slice and the norms.
// Also checks the stash path: gvp_fwd<G, 1> fed with the s' the first evaluation produced.
template <class G, int NS, int ES>
static int run_gvp_split(int n, const float* W, const float* s, const float* v, const float* gs, const float* gv, float* so,
                         float* vo, float* dsin, float* dvin, float* Gacc) {
    static_assert(G::SI == 2 * NS + ES, "split harness: scalar layout");
    int mismatches = 0;
    for (int r = 0; r < n; ++r) {
        float xs[1][G::SI], xv[3][G::VI1], ys[1][G::SO], yv[3][G::VO1];
        for (int k = 0; k < G::SI; ++k) xs[0][k] = s[r * G::SI + k];
        for (int c = 0; c < G::VI; ++c)
            for (int p = 0; p < 3; ++p) xv[p][c] = v[(r * G::VI + c) * 3 + p];
        Save<G> sv;
        // what conv_node_proj_kernel leaves in P: [W_sj s_j ; W_si s_i + b]
        float sj[1][NS], si[1][NS], pj[1][G::SO], pi[1][G::SO];
        for (int k = 0; k < NS; ++k) { sj[0][k] = xs[0][k]; si[0][k] = xs[0][NS + ES + k]; }
        for (int o = 0; o < G::SO; ++o) { pj[0][o] = 0.f; pi[0][o] = W[G::O_WS_T + G::KSD * G::SOP + o]; }
        mv<NS, G::SO, G::SOP, 0, 0>(W + G::O_WS_T, sj, pj);
        mv<NS, G::SO, G::SOP, 0, 0>(W + G::O_WS_T + (NS + ES) * G::SOP, si, pi);
        for (int o = 0; o < G::SO; ++o) sv.sp[0][o] = pj[0][o] + pi[0][o];
        float xe[1][G::SI];
        for (int k = 0; k < G::SI; ++k) xe[0][k] = (k >= NS && k < NS + ES) ? xs[0][k] : 1.0e30f;   // node slices must never be read
        gvp_fwd<G, 2, NS, ES>(W, xe, xv, ys, yv, sv);
        for (int k = 0; k < G::SO; ++k) so[r * G::SO + k] = ys[0][k];
        for (int c = 0; c < G::VO; ++c)
            for (int p = 0; p < 3; ++p) vo[(r * G::VO + c) * 3 + p] = yv[p][c];
        // stash path: same s', everything else recomputed
        {
            Save<G> sv2;
            float y2[1][G::SO], v2[3][G::VO1];
            for (int o = 0; o < G::SO; ++o) sv2.sp[0][o] = sv.sp[0][o];
            gvp_fwd<G, 1>(W, xe, xv, y2, v2, sv2);
            for (int k = 0; k < G::SO; ++k) mismatches += y2[0][k] != ys[0][k];
            for (int c = 0; c < G::VO; ++c)
                for (int p = 0; p < 3; ++p) mismatches += v2[p][c] != yv[p][c];
        }
        float g1[1][G::SO], g3[3][G::VO1], di[1][G::KSD], dv[3][G::VI1], ds[1][G::SO];
        for (int k = 0; k < G::SO; ++k) g1[0][k] = gs[r * G::SO + k];
        for (int c = 0; c < G::VO; ++c)
            for (int p = 0; p < 3; ++p) g3[p][c] = gv[(r * G::VO + c) * 3 + p];
        DirectSink sink{Gacc};
        gvp_bwd_ds<G, DirectSink, true, NS, ES>(W, sv, xe, xv, g1, g3, sink, 0, di, dv, ds);
        // node-level finish (conv_node_post_kernel) for this one row
        for (int k = 0; k < NS; ++k) {
            float a = 0.f, b = 0.f;
            for (int o = 0; o < G::SO; ++o) {
                a += ds[0][o] * W[G::O_WS_B + o * G::KSDP + k];
                b += ds[0][o] * W[G::O_WS_B + o * G::KSDP + NS + ES + k];
                Gacc[G::O_WS_T + k * G::SOP + o] += sj[0][k] * ds[0][o];
                Gacc[G::O_WS_T + (NS + ES + k) * G::SOP + o] += si[0][k] * ds[0][o];
            }
            di[0][k] = a;
            di[0][NS + ES + k] = b;
        }
        for (int o = 0; o < G::SO; ++o) Gacc[G::O_WS_T + G::KSD * G::SOP + o] += ds[0][o];
        for (int k = 0; k < G::SI; ++k) dsin[r * G::SI + k] = di[0][k];
        for (int c = 0; c < G::VI; ++c)
            for (int p = 0; p < 3; ++p) dvin[(r * G::VI + c) * 3 + p] = dv[p][c];
    }
    return mismatches;
}

template <int S, int C>
static void run_ln(int n, const float* w, const float* b, const float* s, const float* v, const float* gs, const float* gv,
                   float* so, float* vo, float* ds, float* dv, float* dw, float* db) {
    for (int r = 0; r < n; ++r) {
        float xs[1][S], xv[3][max1(C)], ys[1][S], yv[3][max1(C)], g1[1][S], g3[3][max1(C)], d1[1][S], d3[3][max1(C)], xh[1][S];
        for (int k = 0; k < S; ++k) { xs[0][k] = s[r * S + k]; g1[0][k] = gs[r * S + k]; }
        for (int c = 0; c < C; ++c)
            for (int p = 0; p < 3; ++p) { xv[p][c] = v[(r * C + c) * 3 + p]; g3[p][c] = gv[(r * C + c) * 3 + p]; }
        const LnStat st = ln_fwd<S, C>(xs, xv, w, b, ys, yv);
        ln_bwd<S, C>(xs, xv, st, w, g1, g3, d1, d3, xh);
        for (int k = 0; k < S; ++k) { so[r * S + k] = ys[0][k]; ds[r * S + k] = d1[0][k]; dw[k] += g1[0][k] * xh[0][k]; db[k] += g1[0][k]; }
        for (int c = 0; c < C; ++c)
            for (int p = 0; p < 3; ++p) { vo[(r * C + c) * 3 + p] = yv[p][c]; dv[(r * C + c) * 3 + p] = d3[p][c]; }
    }
}

#define R CGVP_ACT_RELU
#define N0 CGVP_ACT_NONE
#define SG CGVP_ACT_SIGMOID
extern "C" int harness_gvp(int which, int n, const float* W, const float* s, const float* v, const float* gs, const float* gv,
                           float* so, float* vo, float* dsin, float* dvin, float* G) {
    switch (which) {
        case 0: run_gvp<GvpC<64, 9, 16, 4, 9, R, N0, 1>>(n, W, s, v, gs, gv, so, vo, dsin, dvin, G); return 0;   // message 0
        case 1: run_gvp<GvpC<16, 4, 16, 4, 4, R, N0, 1>>(n, W, s, v, gs, gv, so, vo, dsin, dvin, G); return 0;   // message 1
        case 2: run_gvp<GvpC<16, 4, 16, 4, 4, N0, N0, 1>>(n, W, s, v, gs, gv, so, vo, dsin, dvin, G); return 0;  // message 2
        case 3: run_gvp<GvpC<37, 3, 16, 4, 4, N0, N0, 1>>(n, W, s, v, gs, gv, so, vo, dsin, dvin, G); return 0;  // gvp_node
        case 4: run_gvp<GvpC<33, 1, 32, 1, 1, N0, N0, 1>>(n, W, s, v, gs, gv, so, vo, dsin, dvin, G); return 0;  // gvp_edge
        case 5: run_gvp<GvpC<16, 4, 64, 8, 8, R, N0, 1>>(n, W, s, v, gs, gv, so, vo, dsin, dvin, G); return 0;   // ff 0
        case 6: run_gvp<GvpC<64, 8, 16, 4, 8, N0, N0, 1>>(n, W, s, v, gs, gv, so, vo, dsin, dvin, G); return 0;  // ff 1
        case 7: run_gvp<GvpC<16, 4, 64, 0, 4, R, N0, 1>>(n, W, s, v, gs, gv, so, vo, dsin, dvin, G); return 0;   // gvp_to_scalar
        case 8: run_gvp<GvpC<10, 3, 7, 5, 5, R, SG, 0>>(n, W, s, v, gs, gv, so, vo, dsin, dvin, G); return 0;    // CPD style
        case 9: run_gvp<GvpC<6, 0, 5, 0, 0, R, N0, 0>>(n, W, s, v, gs, gv, so, vo, dsin, dvin, G); return 0;     // scalar only
        case 10: run_gvp<GvpC<12, 5, 9, 3, 6, SG, N0, 0>>(n, W, s, v, gs, gv, so, vo, dsin, dvin, G); return 0;  // no gate, no vact
        // message 0 through the node-projection split + stash path; returns the number of stash-path mismatches
        case 100: return run_gvp_split<GvpC<64, 9, 16, 4, 9, R, N0, 1>, 16, 32>(n, W, s, v, gs, gv, so, vo, dsin, dvin, G);
    }
    return -1;
}
extern "C" int harness_ln(int which, int n, const float* w, const float* b, const float* s, const float* v, const float* gs,
                          const float* gv, float* so, float* vo, float* ds, float* dv, float* dw, float* db) {
    switch (which) {
        case 0: run_ln<16, 4>(n, w, b, s, v, gs, gv, so, vo, ds, dv, dw, db); return 0;
        case 1: run_ln<32, 1>(n, w, b, s, v, gs, gv, so, vo, ds, dv, dw, db); return 0;
        case 2: run_ln<7, 0>(n, w, b, s, v, gs, gv, so, vo, ds, dv, dw, db); return 0;
    }
    return -1;
}

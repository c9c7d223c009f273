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

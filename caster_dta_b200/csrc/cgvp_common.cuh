// Shared host/device definitions for libcastergvp (sm_100a).
//
// Data layout in shared memory ("tile"): every per-row quantity lives in float4 COLUMNS,
//     tile[col4 * RP + row]   (float4),   RP = rows-per-tile + 1,
// i.e. feature-major with a 4-float granule.  A thread that owns row r reads 4 consecutive features of its row
// with one LDS.128; consecutive threads touch consecutive float4s (conflict-free).  The same layout is the
// canonical K-major / no-swizzle operand layout of tcgen05.mma (8x16B core matrices, SBO = 128 B, LBO = RP*16 B).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <type_traits>

#include "../../include/castergvp.h"

#define CGVP_THREADS 128
#define CGVP_EPS 1e-8f
#define CGVP_LN_EPS 1e-5f

// ---- error plumbing -------------------------------------------------------------------------------------------
void cgvp_set_error(const char* fmt, ...);
#define CGVP_REQUIRE(cond, ...)            \
    do {                                   \
        if (!(cond)) {                     \
            cgvp_set_error(__VA_ARGS__);   \
            return -1;                     \
        }                                  \
    } while (0)
#define CGVP_CUDA(call)                                                                  \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            cgvp_set_error("%s failed: %s", #call, cudaGetErrorString(e__));             \
            return (int32_t)e__;                                                         \
        }                                                                                \
    } while (0)
#define CGVP_LAUNCH_CHECK(what)                                                          \
    do {                                                                                 \
        cudaError_t e__ = cudaGetLastError();                                            \
        if (e__ != cudaSuccess) {                                                        \
            cgvp_set_error("launch of %s failed: %s", what, cudaGetErrorString(e__));    \
            return (int32_t)e__;                                                         \
        }                                                                                \
    } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// ---- packed GVP parameters --------------------------------------------------------------------------------------
// Forward half (also the layout of the weight-gradient block), all K-major with the output index contiguous:
//   wh_t  [vi_p][h_p]   wh_t[k][o]  = wh[o][k]
//   ws_t  [ks_p][so_p]  rows k < si+h: ws[o][k];  row k = si+h: bias (the S tile carries a constant-1 column)
//   wv_t  [h_p][vo_p]
//   wsv_t [ksv_p][vo_p] rows k < so: wsv[o][k];   row k = so: gate bias
// Backward half (data gradients; roles of K and N swapped):
//   wh_b [h_p][vi_p]   ws_b [so_p][ksd_p] (ksd = si+h)   wv_b [vo_p][h_p]   wsv_b [vo_p][so_p]
// *_p = padded to a multiple of 4, padding is zero.
struct GvpP {
    int si, vi, so, vo, h;
    int sact, vact, gate;
    int vi4, h4, so4, vo4;   // float4 column counts of vi, h, so, vo
    int ks, ks4;             // ws K = si + h + 1 (ones column for the bias), and its float4 count
    int ksv4;                // gate K = so + 1
    int ksd4;                // si + h in float4 columns (data-gradient N)
    int o_wh_t, o_ws_t, o_wv_t, o_wsv_t, fwd_floats;
    int o_wh_b, o_ws_b, o_wv_b, o_wsv_b, total_floats;
};

static inline GvpP make_gvp_p(const CgvpGvpDesc& d) {
    GvpP g;
    memset(&g, 0, sizeof(g));
    g.si = d.si; g.vi = d.vi; g.so = d.so; g.vo = d.vo;
    g.h = d.vi > 0 ? d.h : 0;
    g.sact = d.scalar_act; g.vact = d.vector_act; g.gate = (d.vector_gate && d.vi > 0 && d.vo > 0) ? 1 : 0;
    g.vi4 = cdiv(g.vi, 4); g.h4 = cdiv(g.h, 4); g.so4 = cdiv(g.so, 4); g.vo4 = cdiv(g.vo, 4);
    g.ks = g.si + g.h + 1; g.ks4 = cdiv(g.ks, 4);
    g.ksv4 = cdiv(g.so + 1, 4);
    g.ksd4 = cdiv(g.si + g.h, 4);
    int o = 0;
    g.o_wh_t = o; o += (g.vi4 * 4) * (g.h4 * 4);
    g.o_ws_t = o; o += (g.ks4 * 4) * (g.so4 * 4);
    g.o_wv_t = o; o += (g.vi > 0 ? (g.h4 * 4) * (g.vo4 * 4) : 0);
    g.o_wsv_t = o; o += (g.gate ? (g.ksv4 * 4) * (g.vo4 * 4) : 0);
    g.fwd_floats = o;
    g.o_wh_b = o; o += (g.h4 * 4) * (g.vi4 * 4);
    g.o_ws_b = o; o += (g.so4 * 4) * (g.ksd4 * 4);
    g.o_wv_b = o; o += (g.vi > 0 ? (g.vo4 * 4) * (g.h4 * 4) : 0);
    g.o_wsv_b = o; o += (g.gate ? (g.vo4 * 4) * (g.so4 * 4) : 0);
    g.total_floats = o;
    return g;
}

int cgvp_validate_gvp(const CgvpGvpDesc& d, const char* what);

// ---- tile column plan for a chain of GVPs ---------------------------------------------------------------------
struct ChainCols {
    int s[CGVP_MAX_CHAIN + 1];    // scalar buffer of stage k (input of GVP k; stage n = chain output)
    int v[CGVP_MAX_CHAIN + 1];    // vector buffer (3 planes, plane pitch vpc[k] columns)
    int vpc[CGVP_MAX_CHAIN + 1];
    int vh[CGVP_MAX_CHAIN], vhpc[CGVP_MAX_CHAIN];   // hidden vectors Vh of GVP k
    int sp[CGVP_MAX_CHAIN];       // gate input (vector_act(s') or s'), with a ones column at index so
    int vo[CGVP_MAX_CHAIN];       // pre-gate output vectors (saved for backward), plane pitch vo4
    int sg[CGVP_MAX_CHAIN];       // gate value sigma (saved for backward)
    int ncols;
};

// distinct = every stage keeps its own buffers (needed by backward); otherwise stages ping-pong.
static inline ChainCols plan_chain_cols(const GvpP* g, int n, bool distinct, bool saves, int start_col) {
    ChainCols c;
    memset(&c, 0, sizeof(c));
    int col = start_col;
    auto scols = [&](int k) { return k < n ? g[k].ks4 : g[n - 1].so4; };
    auto vpcs = [&](int k) { return k < n ? g[k].vi4 : g[n - 1].vo4; };
    if (distinct) {
        for (int k = 0; k <= n; ++k) {
            c.s[k] = col; col += scols(k);
            c.vpc[k] = vpcs(k); c.v[k] = col; col += 3 * c.vpc[k];
        }
        for (int k = 0; k < n; ++k) {
            c.vhpc[k] = g[k].h4; c.vh[k] = col; col += 3 * g[k].h4;
            c.sp[k] = col; col += g[k].ksv4;
            if (saves) {
                c.vo[k] = col; col += 3 * g[k].vo4;
                c.sg[k] = col; col += g[k].vo4;
            }
        }
    } else {
        int smax[2] = {0, 0}, vmax[2] = {0, 0}, vhmax = 0, spmax = 0;
        for (int k = 0; k <= n; ++k) {
            if (scols(k) > smax[k & 1]) smax[k & 1] = scols(k);
            if (3 * vpcs(k) > vmax[k & 1]) vmax[k & 1] = 3 * vpcs(k);
        }
        for (int k = 0; k < n; ++k) {
            if (3 * g[k].h4 > vhmax) vhmax = 3 * g[k].h4;
            if (g[k].ksv4 > spmax) spmax = g[k].ksv4;
        }
        int sreg[2], vreg[2];
        sreg[0] = col; col += smax[0]; vreg[0] = col; col += vmax[0];
        sreg[1] = col; col += smax[1]; vreg[1] = col; col += vmax[1];
        int vhreg = col; col += vhmax;
        int spreg = col; col += spmax;
        for (int k = 0; k <= n; ++k) { c.s[k] = sreg[k & 1]; c.v[k] = vreg[k & 1]; c.vpc[k] = vpcs(k); }
        for (int k = 0; k < n; ++k) { c.vh[k] = vhreg; c.vhpc[k] = g[k].h4; c.sp[k] = spreg; }
    }
    c.ncols = col - start_col;
    return c;
}

// gradient scratch columns shared by all GVPs of a chain (backward)
struct GradCols {
    int gs[2], gv[2], gvpc;   // two (scalar, vector) gradient sets; GVP k reads set (k&1)^par and writes the other
    int dg;                   // gate pre-activation gradient
    int dvh, dvhpc;           // hidden vector gradient
    int ncols;
};

static inline GradCols plan_grad_cols(const GvpP* g, int n, int start_col) {
    GradCols c;
    int smax = 0, vmax = 0, hmax = 0, vomax = 0;
    for (int k = 0; k < n; ++k) {
        if (g[k].ks4 > smax) smax = g[k].ks4;
        if (g[k].so4 > smax) smax = g[k].so4;
        if (g[k].vi4 > vmax) vmax = g[k].vi4;
        if (g[k].vo4 > vmax) vmax = g[k].vo4;
        if (g[k].h4 > hmax) hmax = g[k].h4;
        if (g[k].vo4 > vomax) vomax = g[k].vo4;
    }
    int col = start_col;
    c.gvpc = vmax;
    for (int i = 0; i < 2; ++i) { c.gs[i] = col; col += smax; c.gv[i] = col; col += 3 * vmax; }
    c.dg = col; col += vomax;
    c.dvhpc = hmax; c.dvh = col; col += 3 * hmax;
    c.ncols = col - start_col;
    return c;
}

// optional event bracketing of the main kernels (see cgvp_profile_enable)
void cgvp_prof_begin(int kernel_id, cudaStream_t st);
void cgvp_prof_end(int kernel_id, cudaStream_t st);

int cgvp_max_smem_optin();
int cgvp_num_sms();

// Deterministic reduction of per-CTA partial arenas: reduced[i] = sum_c partial[c*stride + i] in fixed order, then
// segments of `reduced` are copied to their destinations (NULL destinations are skipped).
#define CGVP_MAX_SEGS 8
struct CgvpSeg { float* dst; int off, n; };
int cgvp_reduce_partials(const float* partial, int nparts, int stride, float* reduced, const CgvpSeg* segs, int nsegs,
                         cudaStream_t stream);

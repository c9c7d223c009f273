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
#define CGVP_MAX_THREADS 512   // generic conv kernels at wide dims: several threads per row
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

#define CGVP_HD __host__ __device__
constexpr CGVP_HD inline int cdiv(int a, int b) { return (a + b - 1) / b; }
constexpr CGVP_HD inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }
constexpr CGVP_HD inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
constexpr CGVP_HD inline int imax(int a, int b) { return a > b ? a : b; }
// opt-in dynamic shared memory per block on sm_100 (227 KB); the specialised kernels size their tiles against it
#define CGVP_SMEM_OPTIN 232448

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

constexpr CGVP_HD inline GvpP make_gvp_p(const CgvpGvpDesc& d) {
    GvpP g{};
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
constexpr CGVP_HD inline ChainCols plan_chain_cols(const GvpP* g, int n, bool distinct, bool saves, int start_col) {
    ChainCols c{};
    int col = start_col;
    if (distinct) {
        for (int k = 0; k <= n; ++k) {
            c.s[k] = col; col += k < n ? g[k].ks4 : g[n - 1].so4;
            c.vpc[k] = k < n ? g[k].vi4 : g[n - 1].vo4; c.v[k] = col; col += 3 * c.vpc[k];
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
            const int sc = k < n ? g[k].ks4 : g[n - 1].so4, vc = 3 * (k < n ? g[k].vi4 : g[n - 1].vo4);
            smax[k & 1] = imax(smax[k & 1], sc);
            vmax[k & 1] = imax(vmax[k & 1], vc);
        }
        for (int k = 0; k < n; ++k) {
            vhmax = imax(vhmax, 3 * g[k].h4);
            spmax = imax(spmax, g[k].ksv4);
        }
        int sreg[2] = {0, 0}, vreg[2] = {0, 0};
        sreg[0] = col; col += smax[0]; vreg[0] = col; col += vmax[0];
        sreg[1] = col; col += smax[1]; vreg[1] = col; col += vmax[1];
        const int vhreg = col; col += vhmax;
        const int spreg = col; col += spmax;
        for (int k = 0; k <= n; ++k) { c.s[k] = sreg[k & 1]; c.v[k] = vreg[k & 1]; c.vpc[k] = k < n ? g[k].vi4 : g[n - 1].vo4; }
        for (int k = 0; k < n; ++k) { c.vh[k] = vhreg; c.vhpc[k] = g[k].h4; c.sp[k] = spreg; }
    }
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

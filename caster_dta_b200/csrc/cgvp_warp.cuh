// Warp-level plumbing shared by the register-resident kernels (conv_reg.cu, rows_reg.cu): one warp owns 32
// consecutive rows and never synchronises with the rest of the CTA inside its tile loop.
//   * WarpSink     - weight-gradient GEMMs  G += A^T B  over the warp's 32 rows (shared-memory staged, FFMA2),
//                    accumulated in a warp-private shared-memory arena (deterministic, no atomics)
//   * seg_reduce   - deterministic segmented sum of per-row channels over the (sorted) target node
//   * row loaders  - global <-> register row moves
#pragma once
#include "cgvp_reg.cuh"

namespace cgvpr {

#define CGVP_WPITCH 33   // row pitch of warp-private staging columns (float4 or float): conflict-free transposes

// ---- row moves --------------------------------------------------------------------------------------------------
// dst[0][OFF + i] = base[row * W + i]
template <int W, int OFF, int DN>
__device__ __forceinline__ void load_s(const float* __restrict__ base, long long row, float (&dst)[1][DN]) {
    static_assert(OFF + W <= DN, "load_s: out of range");
    const float* p = base + row * W;
    if constexpr (W % 4 == 0) {
#pragma unroll
        for (int i = 0; i < W / 4; ++i) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
            dst[0][OFF + 4 * i] = t.x; dst[0][OFF + 4 * i + 1] = t.y; dst[0][OFF + 4 * i + 2] = t.z; dst[0][OFF + 4 * i + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < W; ++i) dst[0][OFF + i] = __ldg(p + i);
    }
}
// dst[p][OFF + c] = base[row * 3C + 3c + p]      (rows are [C][3], xyz innermost)
template <int C, int OFF, int DN>
__device__ __forceinline__ void load_v(const float* __restrict__ base, long long row, float (&dst)[3][DN]) {
    static_assert(OFF + C <= DN, "load_v: out of range");
    if constexpr (C > 0) {
        float t[1][3 * C];
        load_s<3 * C, 0>(base, row, t);
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
            for (int p = 0; p < 3; ++p) dst[p][OFF + c] = t[0][3 * c + p];
    }
}
// base[row * W + i] (+)= src[0][OFF + i]
template <int W, int OFF, int SN>
__device__ __forceinline__ void store_s(float* __restrict__ base, long long row, const float (&src)[1][SN], bool accumulate) {
    float* p = base + row * W;
    if constexpr (W % 4 == 0) {
#pragma unroll
        for (int i = 0; i < W / 4; ++i) {
            float4 t = make_float4(src[0][OFF + 4 * i], src[0][OFF + 4 * i + 1], src[0][OFF + 4 * i + 2], src[0][OFF + 4 * i + 3]);
            float4* q = reinterpret_cast<float4*>(p) + i;
            if (accumulate) { const float4 o = *q; t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w; }
            *q = t;
        }
    } else {
#pragma unroll
        for (int i = 0; i < W; ++i) p[i] = accumulate ? p[i] + src[0][OFF + i] : src[0][OFF + i];
    }
}
template <int C, int OFF, int SN>
__device__ __forceinline__ void store_v(float* __restrict__ base, long long row, const float (&src)[3][SN], bool accumulate) {
    if constexpr (C > 0) {
        float t[1][3 * C];
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
            for (int p = 0; p < 3; ++p) t[0][3 * c + p] = src[p][OFF + c];
        store_s<3 * C, 0>(base, row, t, accumulate);
    }
}

// ---- weight-gradient sink ---------------------------------------------------------------------------------------
// Staging layout (float4 columns of 32 rows, pitch 33): A planes first, then B planes.  Lane q owns the 4x4
// blocks q, q+32, ... of the matrix; each block streams the 32 rows (2 LDS.128 + 8 FFMA2 per row) and is then
// added to the warp's arena (plain read-modify-write: the arena is private to the warp).
constexpr CGVP_HD inline int cdiv4(int x) { return (x + 3) / 4; }

template <class G>
constexpr CGVP_HD inline int sink_cols() {
    int m = cdiv4(G::KS) + cdiv4(G::SO);
    if (G::VI > 0) m = imax(m, 3 * (cdiv4(G::VI) + cdiv4(G::H)));
    if (G::VI > 0 && G::VO > 0) m = imax(m, 3 * (cdiv4(G::H) + cdiv4(G::VO)));
    if (G::GATE) m = imax(m, cdiv4(G::KSV) + cdiv4(G::VO));
    return m;
}

struct WarpSink {
    float4* stg;     // warp-private staging, >= cols * CGVP_WPITCH float4
    float* arena;    // warp-private gradient arena
    int lane;
    bool valid;      // this lane's row exists (invalid lanes contribute zeros)

    template <int KA, int NB, int NP, int AX, int BX>
    __device__ __forceinline__ void add(int off, const float (&A)[NP][AX], const float (&B)[NP][BX]) {
        constexpr int KA4 = cdiv4(KA), NB4 = cdiv4(NB);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
#pragma unroll
            for (int a4 = 0; a4 < KA4; ++a4) {
                float4 t;
                t.x = (valid && 4 * a4 + 0 < KA) ? A[p][4 * a4 + 0 < KA ? 4 * a4 + 0 : 0] : 0.f;
                t.y = (valid && 4 * a4 + 1 < KA) ? A[p][4 * a4 + 1 < KA ? 4 * a4 + 1 : 0] : 0.f;
                t.z = (valid && 4 * a4 + 2 < KA) ? A[p][4 * a4 + 2 < KA ? 4 * a4 + 2 : 0] : 0.f;
                t.w = (valid && 4 * a4 + 3 < KA) ? A[p][4 * a4 + 3 < KA ? 4 * a4 + 3 : 0] : 0.f;
                stg[(p * KA4 + a4) * CGVP_WPITCH + lane] = t;
            }
#pragma unroll
            for (int b4 = 0; b4 < NB4; ++b4) {
                float4 t;
                t.x = (valid && 4 * b4 + 0 < NB) ? B[p][4 * b4 + 0 < NB ? 4 * b4 + 0 : 0] : 0.f;
                t.y = (valid && 4 * b4 + 1 < NB) ? B[p][4 * b4 + 1 < NB ? 4 * b4 + 1 : 0] : 0.f;
                t.z = (valid && 4 * b4 + 2 < NB) ? B[p][4 * b4 + 2 < NB ? 4 * b4 + 2 : 0] : 0.f;
                t.w = (valid && 4 * b4 + 3 < NB) ? B[p][4 * b4 + 3 < NB ? 4 * b4 + 3 : 0] : 0.f;
                stg[(NP * KA4 + p * NB4 + b4) * CGVP_WPITCH + lane] = t;
            }
        }
        __syncwarp();
        constexpr int NBLK = KA4 * NB4;
        // full rounds: lane q owns block q of the round and streams all 32 rows
#pragma unroll 1
        for (int b0 = 0; b0 + 32 <= NBLK; b0 += 32) block_rows<KA4, NB4, NP, 32, 0>(off, b0 + lane, 0);
        // remainder (< 32 blocks): split the 32 rows over G lane groups, then a fixed-order butterfly over the groups
        constexpr int REM = NBLK % 32;
        if constexpr (REM > 0) {
            constexpr int P = REM <= 1 ? 1 : (REM <= 2 ? 2 : (REM <= 4 ? 4 : (REM <= 8 ? 8 : (REM <= 16 ? 16 : 32))));
            constexpr int G = 32 / P;
            const int blk = lane % P, grp = lane / P;
            if (G == 1) { if (blk < REM) block_rows<KA4, NB4, NP, 32, 0>(off, NBLK - REM + blk, 0); }
            else block_rows<KA4, NB4, NP, 32 / G, G>(off, NBLK - REM + (blk < REM ? blk : 0), grp * (32 / G), blk < REM && grp == 0);
        }
        __syncwarp();
    }

    // One 4x4 block over ROWS staged rows starting at r0.  G > 1: the partial blocks of the G lane groups are summed
    // with G-1 shuffle exchanges (xor butterfly, fixed order) and lanes with `commit` add the result to the arena.
    template <int KA4, int NB4, int NP, int ROWS, int G>
    __device__ __forceinline__ void block_rows(int off, int blk, int r0, bool commit = true) {
        constexpr int NBP = NB4 * 4;
        const int a4 = blk / NB4, b4 = blk - a4 * NB4;
        float2 c[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i] = make_float2(0.f, 0.f);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const float4* Ap = stg + (p * KA4 + a4) * CGVP_WPITCH + r0;
            const float4* Bp = stg + (NP * KA4 + p * NB4 + b4) * CGVP_WPITCH + r0;
#pragma unroll(ROWS < 8 ? ROWS : 8)
            for (int r = 0; r < ROWS; ++r) {
                const float4 x = Ap[r], y = Bp[r];
                fma2(c[0], x.x, make_float2(y.x, y.y)); fma2(c[1], x.x, make_float2(y.z, y.w));
                fma2(c[2], x.y, make_float2(y.x, y.y)); fma2(c[3], x.y, make_float2(y.z, y.w));
                fma2(c[4], x.z, make_float2(y.x, y.y)); fma2(c[5], x.z, make_float2(y.z, y.w));
                fma2(c[6], x.w, make_float2(y.x, y.y)); fma2(c[7], x.w, make_float2(y.z, y.w));
            }
        }
        if constexpr (G > 1) {
#pragma unroll
            for (int o = 32 / G; o < 32; o <<= 1) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    c[i].x += __shfl_xor_sync(0xffffffffu, c[i].x, o);
                    c[i].y += __shfl_xor_sync(0xffffffffu, c[i].y, o);
                }
            }
        }
        if (commit) {
            float* g = arena + off + (a4 * 4) * NBP + b4 * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4* q = reinterpret_cast<float4*>(g + i * NBP);
                float4 o = *q;
                o.x += c[2 * i].x; o.y += c[2 * i].y; o.z += c[2 * i + 1].x; o.w += c[2 * i + 1].y;
                *q = o;
            }
        }
    }
};

// ---- segmented reduction of the warp's rows over the sorted target node -----------------------------------------
// M[ch * CGVP_WPITCH + r]: channel ch of row r;  dsts[r]: target node of row r (non-decreasing).
// A segment that lies completely inside the tile is written to out (x 1/deg for mean); a segment that started in
// an earlier tile goes to part_head[tile], one that continues into a later tile to part_tail[tile]; the fix-up
// kernel adds those pieces in tile order.  Channels ch < SW go to out_s[n * SW + ch], the rest to out_v.
template <int CH, int SW>
__device__ __forceinline__ void seg_reduce_warp(const float* M, const int* dsts, int lane, int rv, long long p0, int tile,
                                                const int* __restrict__ rowptr, bool mean, float* __restrict__ out_s,
                                                float* __restrict__ out_v, float* __restrict__ part_head,
                                                float* __restrict__ part_tail) {
    const long long p1 = p0 + rv;
    for (int ch = lane; ch < CH; ch += 32) {
        int cur = dsts[0];
        float sum = 0.f;
        for (int r = 0; r <= rv; ++r) {
            const int n = r < rv ? dsts[r] : -1;
            if (n != cur) {
                const long long a = rowptr[cur], b = rowptr[cur + 1];
                if (a >= p0 && b <= p1) {
                    const float f = mean ? 1.f / (float)max((int)(b - a), 1) : 1.f;
                    if (ch < SW) out_s[(long long)cur * SW + ch] = sum * f;
                    else out_v[(long long)cur * (CH - SW) + (ch - SW)] = sum * f;
                } else if (a < p0) {
                    part_head[(long long)tile * CH + ch] = sum;
                } else {
                    part_tail[(long long)tile * CH + ch] = sum;
                }
                cur = n;
                sum = 0.f;
            }
            if (r < rv) sum += M[ch * CGVP_WPITCH + r];
        }
    }
}

}  // namespace cgvpr

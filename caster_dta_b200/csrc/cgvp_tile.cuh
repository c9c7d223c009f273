// Device-side building blocks: one thread owns one row (edge or node) of a shared-memory tile and runs the whole
// GVP chain on it with packed fp32x2 FMAs (FFMA2, sm_100a); weights are read at warp-uniform addresses.
// Restates the maths of models/gvp_layers.py:142-175 (GVP), :231-242 (LayerNorm) and SURVEY.md Appendix E (backward).
#pragma once
#include "cgvp_common.cuh"

namespace cgvp {

// ---- scalar helpers -------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_fwd(int a, float x) {
    if (a == CGVP_ACT_RELU) return fmaxf(x, 0.f);
    if (a == CGVP_ACT_SIGMOID) return 1.f / (1.f + expf(-x));
    return x;
}
// derivative of the activation expressed through its OUTPUT y
__device__ __forceinline__ float act_bwd(int a, float y) {
    if (a == CGVP_ACT_RELU) return y > 0.f ? 1.f : 0.f;
    if (a == CGVP_ACT_SIGMOID) return y * (1.f - y);
    return 1.f;
}
__device__ __forceinline__ float4 act_fwd4(int a, float4 x) {
    return make_float4(act_fwd(a, x.x), act_fwd(a, x.y), act_fwd(a, x.z), act_fwd(a, x.w));
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// c += a * b  on two lanes at once (FFMA2; the scalar multiplicand is broadcast by the hardware operand form)
__device__ __forceinline__ void fma2(float2& c, float a, float2 b) {
    unsigned long long cc, aa, bb;
    asm("mov.b64 %0, {%1, %2};" : "=l"(cc) : "f"(c.x), "f"(c.y));
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b.x), "f"(b.y));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(cc) : "l"(aa), "l"(bb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(c.x), "=f"(c.y) : "l"(cc));
}

__device__ __forceinline__ float f4get(const float4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

// ---- the row-linear primitive ---------------------------------------------------------------------------------
// acc[p][0..OB) += sum_k in_p[k] * W[k*ldw + o0 + 0..OB)   for NP planes sharing the weights.
// `in` points at (column 0, this thread's row); column stride = rp float4; plane stride = pp float4.
template <int OB, int NP>
__device__ __forceinline__ void rl_block(const float4* __restrict__ in, int pp, int rp, int k4n,
                                         const float* __restrict__ W, int ldw, int o0, float2 (&acc)[NP][OB / 2]) {
    const float* __restrict__ w = W + o0;
#pragma unroll
    for (int k4 = 0; k4 < k4n; ++k4) {
        float4 x[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) x[p] = in[p * pp + k4 * rp];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int j = 0; j < OB / 4; ++j) {
                const float4 w4 = *reinterpret_cast<const float4*>(w + (k4 * 4 + kk) * ldw + 4 * j);
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const float xv = f4get(x[p], kk);
                    fma2(acc[p][2 * j], xv, make_float2(w4.x, w4.y));
                    fma2(acc[p][2 * j + 1], xv, make_float2(w4.z, w4.w));
                }
            }
        }
    }
}

template <int N>
struct IC { static constexpr int value = N; };

// split n4 float4-columns of outputs into register blocks of 16 / 8 / 4 outputs
template <class F>
__device__ __forceinline__ void for_blocks(int n4, F&& f) {
    int o4 = 0;
#pragma unroll
    for (; n4 - o4 >= 4; o4 += 4) f(IC<16>{}, o4 * 4);
    if (n4 - o4 >= 2) { f(IC<8>{}, o4 * 4); o4 += 2; }
    if (n4 - o4 >= 1) f(IC<4>{}, o4 * 4);
}

template <int NP, int H>
__device__ __forceinline__ void zero_acc(float2 (&a)[NP][H]) {
#pragma unroll
    for (int p = 0; p < NP; ++p)
#pragma unroll
        for (int i = 0; i < H; ++i) a[p][i] = make_float2(0.f, 0.f);
}

// column map of one GVP application inside the tile (float4 column indices)
struct StageIO {
    int s_in, v_in, v_in_pc;     // inputs  (S has room for [s ; vn ; 1])
    int vh, vh_pc;               // hidden vectors
    int sp;                      // gate input with ones column
    int s_out, v_out, v_out_pc;  // outputs (= inputs of the next stage)
    int vo, sg;                  // saves for backward: pre-gate V', gate value
};

constexpr CGVP_HD inline StageIO stage_io(const ChainCols& c, int k) {
    StageIO s{};
    s.s_in = c.s[k]; s.v_in = c.v[k]; s.v_in_pc = c.vpc[k];
    s.vh = c.vh[k]; s.vh_pc = c.vhpc[k]; s.sp = c.sp[k];
    s.s_out = c.s[k + 1]; s.v_out = c.v[k + 1]; s.v_out_pc = c.vpc[k + 1];
    s.vo = c.vo[k]; s.sg = c.sg[k];
    return s;
}

__device__ __forceinline__ float& tile_at(float4* T, int rp, int r, int col4base, int col) {
    return reinterpret_cast<float*>(T)[((col4base + (col >> 2)) * rp + r) * 4 + (col & 3)];
}

// ---- GVP forward for one row (gvp_layers.py:142-175) ----------------------------------------------------------
template <bool SAVE>
__device__ __forceinline__ void gvp_fwd_row(const GvpP& g, const float* __restrict__ W, float4* T, int rp, int r,
                                            const StageIO& c) {
    if (g.vi > 0) {
        const int hp = g.h4 * 4;
        for_blocks(g.h4, [&](auto obc, int o0) {
            constexpr int OB = decltype(obc)::value;
            float2 acc[3][OB / 2];
            zero_acc(acc);
            rl_block<OB, 3>(T + c.v_in * rp + r, c.v_in_pc * rp, rp, g.vi4, W + g.o_wh_t, hp, o0, acc);   // :152
#pragma unroll
            for (int j = 0; j < OB / 4; ++j)
#pragma unroll
                for (int p = 0; p < 3; ++p)
                    T[(c.vh + p * c.vh_pc + (o0 >> 2) + j) * rp + r] =
                        make_float4(acc[p][2 * j].x, acc[p][2 * j].y, acc[p][2 * j + 1].x, acc[p][2 * j + 1].y);
#pragma unroll
            for (int i = 0; i < OB / 2; ++i) {                                                              // :153
                const float qx = acc[0][i].x * acc[0][i].x + acc[1][i].x * acc[1][i].x + acc[2][i].x * acc[2][i].x;
                const float qy = acc[0][i].y * acc[0][i].y + acc[1][i].y * acc[1][i].y + acc[2][i].y * acc[2][i].y;
                const int o = o0 + 2 * i;
                if (o < g.h) tile_at(T, rp, r, c.s_in, g.si + o) = sqrtf(fmaxf(qx, CGVP_EPS));
                if (o + 1 < g.h) tile_at(T, rp, r, c.s_in, g.si + o + 1) = sqrtf(fmaxf(qy, CGVP_EPS));
            }
        });
    }
    {   // constant-1 column (bias row of ws_t) and zero padding up to the float4 boundary
        int k = g.si + g.h;
        tile_at(T, rp, r, c.s_in, k) = 1.f;
        for (++k; k < g.ks4 * 4; ++k) tile_at(T, rp, r, c.s_in, k) = 0.f;
    }
    const int sop = g.so4 * 4;
    for_blocks(g.so4, [&](auto obc, int o0) {                                                               // :154
        constexpr int OB = decltype(obc)::value;
        float2 acc[1][OB / 2];
        zero_acc(acc);
        rl_block<OB, 1>(T + c.s_in * rp + r, 0, rp, g.ks4, W + g.o_ws_t, sop, o0, acc);
#pragma unroll
        for (int j = 0; j < OB / 4; ++j) {
            const float4 pre = make_float4(acc[0][2 * j].x, acc[0][2 * j].y, acc[0][2 * j + 1].x, acc[0][2 * j + 1].y);
            T[(c.s_out + (o0 >> 2) + j) * rp + r] = act_fwd4(g.sact, pre);                                 // :172-173
            if (g.gate) T[(c.sp + (o0 >> 2) + j) * rp + r] = act_fwd4(g.vact, pre);                        // :159-162
        }
    });
    if (g.vo > 0) {
        if (g.vi > 0) {
            if (g.gate) {
                tile_at(T, rp, r, c.sp, g.so) = 1.f;
                for (int k = g.so + 1; k < g.ksv4 * 4; ++k) tile_at(T, rp, r, c.sp, k) = 0.f;
            }
            const int vop = g.vo4 * 4;
            for_blocks(g.vo4, [&](auto obc, int o0) {
                constexpr int OB = decltype(obc)::value;
                float2 av[3][OB / 2];
                zero_acc(av);
                rl_block<OB, 3>(T + c.vh * rp + r, c.vh_pc * rp, rp, g.h4, W + g.o_wv_t, vop, o0, av);     // :156
                float2 sg[1][OB / 2];
                if (g.gate) {                                                                              // :158-163
                    zero_acc(sg);
                    rl_block<OB, 1>(T + c.sp * rp + r, 0, rp, g.ksv4, W + g.o_wsv_t, vop, o0, sg);
#pragma unroll
                    for (int i = 0; i < OB / 2; ++i) sg[0][i] = make_float2(sigmoidf_(sg[0][i].x), sigmoidf_(sg[0][i].y));
                } else {
#pragma unroll
                    for (int i = 0; i < OB / 2; ++i) {
                        if (g.vact) {                                                                      // :164-166
                            const float qx = av[0][i].x * av[0][i].x + av[1][i].x * av[1][i].x + av[2][i].x * av[2][i].x;
                            const float qy = av[0][i].y * av[0][i].y + av[1][i].y * av[1][i].y + av[2][i].y * av[2][i].y;
                            sg[0][i] = make_float2(act_fwd(g.vact, sqrtf(fmaxf(qx, CGVP_EPS))),
                                                   act_fwd(g.vact, sqrtf(fmaxf(qy, CGVP_EPS))));
                        } else {
                            sg[0][i] = make_float2(1.f, 1.f);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < OB / 4; ++j) {
                    const float4 s4 = make_float4(sg[0][2 * j].x, sg[0][2 * j].y, sg[0][2 * j + 1].x, sg[0][2 * j + 1].y);
                    if (SAVE) T[(c.sg + (o0 >> 2) + j) * rp + r] = s4;
#pragma unroll
                    for (int p = 0; p < 3; ++p) {
                        const float4 a4 = make_float4(av[p][2 * j].x, av[p][2 * j].y, av[p][2 * j + 1].x, av[p][2 * j + 1].y);
                        if (SAVE) T[(c.vo + p * g.vo4 + (o0 >> 2) + j) * rp + r] = a4;
                        T[(c.v_out + p * c.v_out_pc + (o0 >> 2) + j) * rp + r] =
                            make_float4(a4.x * s4.x, a4.y * s4.y, a4.z * s4.z, a4.w * s4.w);
                    }
                }
            });
        } else {                                                                                           // :169-171
            for (int j = 0; j < 3 * g.vo4; ++j) T[(c.v_out + j) * rp + r] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

// ---- GVP backward for one row ----------------------------------------------------------
// On entry  (gs_in [so], gv_in [3][vo]) hold the gradient of the GVP outputs.
// On exit   they hold ds' and dVo (operands of the weight-gradient GEMMs), `dg` holds the gate pre-activation
//           gradient, `dvh` the hidden-vector gradient, and (gs_out [si+h], gv_out [3][vi]) the input gradients.
struct GradIO {
    int gs_in, gv_in, gv_in_pc;
    int gs_out, gv_out, gv_out_pc;
    int dg, dvh, dvh_pc;
};

__device__ __forceinline__ void gvp_bwd_row(const GvpP& g, const float* __restrict__ W, float4* T, int rp, int r,
                                            const StageIO& c, const GradIO& d) {
    const bool has_v = g.vi > 0 && g.vo > 0;
    if (has_v) {
        for (int o4 = 0; o4 < g.vo4; ++o4) {
            float4 dv[3], vo[3];
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                dv[p] = T[(d.gv_in + p * d.gv_in_pc + o4) * rp + r];
                vo[p] = T[(c.vo + p * g.vo4 + o4) * rp + r];
            }
            const float4 sg = T[(c.sg + o4) * rp + r];
            float4 dgv = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float s = f4get(sg, i);
                const float dx = f4get(dv[0], i), dy = f4get(dv[1], i), dz = f4get(dv[2], i);
                const float vx = f4get(vo[0], i), vy = f4get(vo[1], i), vz = f4get(vo[2], i);
                const float dot = dx * vx + dy * vy + dz * vz;
                float ox, oy, oz, gg = 0.f;
                if (g.gate) {
                    gg = dot * s * (1.f - s);
                    ox = dx * s; oy = dy * s; oz = dz * s;
                } else if (g.vact) {
                    const float q = vx * vx + vy * vy + vz * vz;
                    const float t = q >= CGVP_EPS ? dot * act_bwd(g.vact, s) / sqrtf(q) : 0.f;
                    ox = dx * s + vx * t; oy = dy * s + vy * t; oz = dz * s + vz * t;
                } else {
                    ox = dx; oy = dy; oz = dz;
                }
                reinterpret_cast<float*>(&dv[0])[i] = ox;
                reinterpret_cast<float*>(&dv[1])[i] = oy;
                reinterpret_cast<float*>(&dv[2])[i] = oz;
                reinterpret_cast<float*>(&dgv)[i] = gg;
            }
#pragma unroll
            for (int p = 0; p < 3; ++p) T[(d.gv_in + p * d.gv_in_pc + o4) * rp + r] = dv[p];
            if (g.gate) T[(d.dg + o4) * rp + r] = dgv;
        }
    }
    // ds' = dS_out * sact'(s_out) + (dg . wsv) * vact'(gate_in)
    if (g.gate) {
        const int sop = g.so4 * 4;
        for_blocks(g.so4, [&](auto obc, int o0) {
            constexpr int OB = decltype(obc)::value;
            float2 acc[1][OB / 2];
            zero_acc(acc);
            rl_block<OB, 1>(T + d.dg * rp + r, 0, rp, g.vo4, W + g.o_wsv_b, sop, o0, acc);
#pragma unroll
            for (int j = 0; j < OB / 4; ++j) {
                const int col = (o0 >> 2) + j;
                const float4 ds = T[(d.gs_in + col) * rp + r];
                const float4 so = T[(c.s_out + col) * rp + r];
                const float4 gi = T[(c.sp + col) * rp + r];
                const float a[4] = {acc[0][2 * j].x, acc[0][2 * j].y, acc[0][2 * j + 1].x, acc[0][2 * j + 1].y};
                float4 o;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    reinterpret_cast<float*>(&o)[i] =
                        f4get(ds, i) * act_bwd(g.sact, f4get(so, i)) + a[i] * act_bwd(g.vact, f4get(gi, i));
                T[(d.gs_in + col) * rp + r] = o;
            }
        });
    } else if (g.sact) {
        for (int col = 0; col < g.so4; ++col) {
            float4 ds = T[(d.gs_in + col) * rp + r];
            const float4 so = T[(c.s_out + col) * rp + r];
            ds.x *= act_bwd(g.sact, so.x); ds.y *= act_bwd(g.sact, so.y);
            ds.z *= act_bwd(g.sact, so.z); ds.w *= act_bwd(g.sact, so.w);
            T[(d.gs_in + col) * rp + r] = ds;
        }
    }
    // [dS_in ; dvn] = ds' . ws
    {
        const int np_ = g.ksd4 * 4;
        for_blocks(g.ksd4, [&](auto obc, int o0) {
            constexpr int OB = decltype(obc)::value;
            float2 acc[1][OB / 2];
            zero_acc(acc);
            rl_block<OB, 1>(T + d.gs_in * rp + r, 0, rp, g.so4, W + g.o_ws_b, np_, o0, acc);
#pragma unroll
            for (int j = 0; j < OB / 4; ++j)
                T[(d.gs_out + (o0 >> 2) + j) * rp + r] =
                    make_float4(acc[0][2 * j].x, acc[0][2 * j].y, acc[0][2 * j + 1].x, acc[0][2 * j + 1].y);
        });
    }
    if (g.vi > 0) {
        // dVh = wv^T dVo + Vh * dvn / vn   (clamp passes the gradient where |Vh|^2 >= eps)
        const int hp = g.h4 * 4;
        for_blocks(g.h4, [&](auto obc, int o0) {
            constexpr int OB = decltype(obc)::value;
            float2 acc[3][OB / 2];
            zero_acc(acc);
            if (g.vo > 0) rl_block<OB, 3>(T + d.gv_in * rp + r, d.gv_in_pc * rp, rp, g.vo4, W + g.o_wv_b, hp, o0, acc);
#pragma unroll
            for (int j = 0; j < OB / 4; ++j) {
                const int col = (o0 >> 2) + j;
                float4 vh[3], o[3];
#pragma unroll
                for (int p = 0; p < 3; ++p) vh[p] = T[(c.vh + p * c.vh_pc + col) * rp + r];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int k = o0 + 4 * j + i;
                    const float x = f4get(vh[0], i), y = f4get(vh[1], i), z = f4get(vh[2], i);
                    const float q = x * x + y * y + z * z;
                    float f = 0.f;
                    if (k < g.h && q >= CGVP_EPS)
                        f = tile_at(T, rp, r, d.gs_out, g.si + k) / tile_at(T, rp, r, c.s_in, g.si + k);
                    const float2 a0 = acc[0][2 * j + (i >> 1)], a1 = acc[1][2 * j + (i >> 1)], a2 = acc[2][2 * j + (i >> 1)];
                    reinterpret_cast<float*>(&o[0])[i] = ((i & 1) ? a0.y : a0.x) + x * f;
                    reinterpret_cast<float*>(&o[1])[i] = ((i & 1) ? a1.y : a1.x) + y * f;
                    reinterpret_cast<float*>(&o[2])[i] = ((i & 1) ? a2.y : a2.x) + z * f;
                }
#pragma unroll
                for (int p = 0; p < 3; ++p) T[(d.dvh + p * d.dvh_pc + col) * rp + r] = o[p];
            }
        });
        // dV_in = wh^T dVh
        const int vip = g.vi4 * 4;
        for_blocks(g.vi4, [&](auto obc, int o0) {
            constexpr int OB = decltype(obc)::value;
            float2 acc[3][OB / 2];
            zero_acc(acc);
            rl_block<OB, 3>(T + d.dvh * rp + r, d.dvh_pc * rp, rp, g.h4, W + g.o_wh_b, vip, o0, acc);
#pragma unroll
            for (int j = 0; j < OB / 4; ++j)
#pragma unroll
                for (int p = 0; p < 3; ++p)
                    T[(d.gv_out + p * d.gv_out_pc + (o0 >> 2) + j) * rp + r] =
                        make_float4(acc[p][2 * j].x, acc[p][2 * j].y, acc[p][2 * j + 1].x, acc[p][2 * j + 1].y);
        });
    }
}

// ---- team versions: T threads share one row ------------------------------------------------------------------------
// At wide dims only a few rows fit in the shared-memory tile; a thread-per-row kernel would leave most of the CTA idle.
// Here `tsz` threads (rank 0..tsz-1) work on the SAME row, each taking every tsz-th 8-column output block of a phase;
// phases are separated by CTA barriers (every thread of the CTA must call these functions, `work` = this thread has a
// row).  tsz = 1 reproduces the per-row functions above.
template <class F>
__device__ __forceinline__ void for_blocks_team(int n4, int rank, int tsz, F&& f) {
    const int nb = (n4 + 1) >> 1;                       // blocks of 8 columns, the last one 4 if n4 is odd
    for (int b = rank; b < nb; b += tsz) {
        if (2 * b + 1 < n4) f(IC<8>{}, b * 8);
        else f(IC<4>{}, b * 8);
    }
}

template <bool SAVE>
__device__ __forceinline__ void gvp_fwd_team(const GvpP& g, const float* __restrict__ W, float4* T, int rp, int r,
                                             const StageIO& c, int rank, int tsz, bool work) {
    if (work && g.vi > 0) {
        const int hp = g.h4 * 4;
        for_blocks_team(g.h4, rank, tsz, [&](auto obc, int o0) {
            constexpr int OB = decltype(obc)::value;
            float2 acc[3][OB / 2];
            zero_acc(acc);
            rl_block<OB, 3>(T + c.v_in * rp + r, c.v_in_pc * rp, rp, g.vi4, W + g.o_wh_t, hp, o0, acc);   // :152
#pragma unroll
            for (int j = 0; j < OB / 4; ++j)
#pragma unroll
                for (int p = 0; p < 3; ++p)
                    T[(c.vh + p * c.vh_pc + (o0 >> 2) + j) * rp + r] =
                        make_float4(acc[p][2 * j].x, acc[p][2 * j].y, acc[p][2 * j + 1].x, acc[p][2 * j + 1].y);
#pragma unroll
            for (int i = 0; i < OB / 2; ++i) {                                                              // :153
                const float qx = acc[0][i].x * acc[0][i].x + acc[1][i].x * acc[1][i].x + acc[2][i].x * acc[2][i].x;
                const float qy = acc[0][i].y * acc[0][i].y + acc[1][i].y * acc[1][i].y + acc[2][i].y * acc[2][i].y;
                const int o = o0 + 2 * i;
                if (o < g.h) tile_at(T, rp, r, c.s_in, g.si + o) = sqrtf(fmaxf(qx, CGVP_EPS));
                if (o + 1 < g.h) tile_at(T, rp, r, c.s_in, g.si + o + 1) = sqrtf(fmaxf(qy, CGVP_EPS));
            }
        });
    }
    if (work && rank == 0) {   // constant-1 column (bias row of ws_t) and zero padding up to the float4 boundary
        int k = g.si + g.h;
        tile_at(T, rp, r, c.s_in, k) = 1.f;
        for (++k; k < g.ks4 * 4; ++k) tile_at(T, rp, r, c.s_in, k) = 0.f;
    }
    __syncthreads();
    if (work) {
        const int sop = g.so4 * 4;
        for_blocks_team(g.so4, rank, tsz, [&](auto obc, int o0) {                                           // :154
            constexpr int OB = decltype(obc)::value;
            float2 acc[1][OB / 2];
            zero_acc(acc);
            rl_block<OB, 1>(T + c.s_in * rp + r, 0, rp, g.ks4, W + g.o_ws_t, sop, o0, acc);
#pragma unroll
            for (int j = 0; j < OB / 4; ++j) {
                const float4 pre = make_float4(acc[0][2 * j].x, acc[0][2 * j].y, acc[0][2 * j + 1].x, acc[0][2 * j + 1].y);
                T[(c.s_out + (o0 >> 2) + j) * rp + r] = act_fwd4(g.sact, pre);                             // :172-173
                if (g.gate) T[(c.sp + (o0 >> 2) + j) * rp + r] = act_fwd4(g.vact, pre);                    // :159-162
            }
        });
    }
    __syncthreads();
    if (g.vo > 0 && g.vi > 0 && g.gate) {
        if (work && rank == 0) {
            tile_at(T, rp, r, c.sp, g.so) = 1.f;
            for (int k = g.so + 1; k < g.ksv4 * 4; ++k) tile_at(T, rp, r, c.sp, k) = 0.f;
        }
        __syncthreads();
    }
    if (work && g.vo > 0) {
        if (g.vi > 0) {
            const int vop = g.vo4 * 4;
            for_blocks_team(g.vo4, rank, tsz, [&](auto obc, int o0) {
                constexpr int OB = decltype(obc)::value;
                float2 av[3][OB / 2];
                zero_acc(av);
                rl_block<OB, 3>(T + c.vh * rp + r, c.vh_pc * rp, rp, g.h4, W + g.o_wv_t, vop, o0, av);     // :156
                float2 sg[1][OB / 2];
                if (g.gate) {                                                                              // :158-163
                    zero_acc(sg);
                    rl_block<OB, 1>(T + c.sp * rp + r, 0, rp, g.ksv4, W + g.o_wsv_t, vop, o0, sg);
#pragma unroll
                    for (int i = 0; i < OB / 2; ++i) sg[0][i] = make_float2(sigmoidf_(sg[0][i].x), sigmoidf_(sg[0][i].y));
                } else {
#pragma unroll
                    for (int i = 0; i < OB / 2; ++i) {
                        if (g.vact) {                                                                      // :164-166
                            const float qx = av[0][i].x * av[0][i].x + av[1][i].x * av[1][i].x + av[2][i].x * av[2][i].x;
                            const float qy = av[0][i].y * av[0][i].y + av[1][i].y * av[1][i].y + av[2][i].y * av[2][i].y;
                            sg[0][i] = make_float2(act_fwd(g.vact, sqrtf(fmaxf(qx, CGVP_EPS))),
                                                   act_fwd(g.vact, sqrtf(fmaxf(qy, CGVP_EPS))));
                        } else {
                            sg[0][i] = make_float2(1.f, 1.f);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < OB / 4; ++j) {
                    const float4 s4 = make_float4(sg[0][2 * j].x, sg[0][2 * j].y, sg[0][2 * j + 1].x, sg[0][2 * j + 1].y);
                    if (SAVE) T[(c.sg + (o0 >> 2) + j) * rp + r] = s4;
#pragma unroll
                    for (int p = 0; p < 3; ++p) {
                        const float4 a4 = make_float4(av[p][2 * j].x, av[p][2 * j].y, av[p][2 * j + 1].x, av[p][2 * j + 1].y);
                        if (SAVE) T[(c.vo + p * g.vo4 + (o0 >> 2) + j) * rp + r] = a4;
                        T[(c.v_out + p * c.v_out_pc + (o0 >> 2) + j) * rp + r] =
                            make_float4(a4.x * s4.x, a4.y * s4.y, a4.z * s4.z, a4.w * s4.w);
                    }
                }
            });
        } else {                                                                                           // :169-171
            for (int j = rank; j < 3 * g.vo4; j += tsz) T[(c.v_out + j) * rp + r] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void gvp_bwd_team(const GvpP& g, const float* __restrict__ W, float4* T, int rp, int r,
                                             const StageIO& c, const GradIO& d, int rank, int tsz, bool work) {
    const bool has_v = g.vi > 0 && g.vo > 0;
    if (work && has_v) {
        for (int o4 = rank; o4 < g.vo4; o4 += tsz) {
            float4 dv[3], vo[3];
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                dv[p] = T[(d.gv_in + p * d.gv_in_pc + o4) * rp + r];
                vo[p] = T[(c.vo + p * g.vo4 + o4) * rp + r];
            }
            const float4 sg = T[(c.sg + o4) * rp + r];
            float4 dgv = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float s = f4get(sg, i);
                const float dx = f4get(dv[0], i), dy = f4get(dv[1], i), dz = f4get(dv[2], i);
                const float vx = f4get(vo[0], i), vy = f4get(vo[1], i), vz = f4get(vo[2], i);
                const float dot = dx * vx + dy * vy + dz * vz;
                float ox, oy, oz, gg = 0.f;
                if (g.gate) {
                    gg = dot * s * (1.f - s);
                    ox = dx * s; oy = dy * s; oz = dz * s;
                } else if (g.vact) {
                    const float q = vx * vx + vy * vy + vz * vz;
                    const float t = q >= CGVP_EPS ? dot * act_bwd(g.vact, s) / sqrtf(q) : 0.f;
                    ox = dx * s + vx * t; oy = dy * s + vy * t; oz = dz * s + vz * t;
                } else {
                    ox = dx; oy = dy; oz = dz;
                }
                reinterpret_cast<float*>(&dv[0])[i] = ox;
                reinterpret_cast<float*>(&dv[1])[i] = oy;
                reinterpret_cast<float*>(&dv[2])[i] = oz;
                reinterpret_cast<float*>(&dgv)[i] = gg;
            }
#pragma unroll
            for (int p = 0; p < 3; ++p) T[(d.gv_in + p * d.gv_in_pc + o4) * rp + r] = dv[p];
            if (g.gate) T[(d.dg + o4) * rp + r] = dgv;
        }
    }
    __syncthreads();
    if (work) {   // ds' = dS_out * sact'(s_out) + (dg . wsv) * vact'(gate_in)
        if (g.gate) {
            const int sop = g.so4 * 4;
            for_blocks_team(g.so4, rank, tsz, [&](auto obc, int o0) {
                constexpr int OB = decltype(obc)::value;
                float2 acc[1][OB / 2];
                zero_acc(acc);
                rl_block<OB, 1>(T + d.dg * rp + r, 0, rp, g.vo4, W + g.o_wsv_b, sop, o0, acc);
#pragma unroll
                for (int j = 0; j < OB / 4; ++j) {
                    const int col = (o0 >> 2) + j;
                    const float4 ds = T[(d.gs_in + col) * rp + r];
                    const float4 so = T[(c.s_out + col) * rp + r];
                    const float4 gi = T[(c.sp + col) * rp + r];
                    const float a[4] = {acc[0][2 * j].x, acc[0][2 * j].y, acc[0][2 * j + 1].x, acc[0][2 * j + 1].y};
                    float4 o;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        reinterpret_cast<float*>(&o)[i] =
                            f4get(ds, i) * act_bwd(g.sact, f4get(so, i)) + a[i] * act_bwd(g.vact, f4get(gi, i));
                    T[(d.gs_in + col) * rp + r] = o;
                }
            });
        } else if (g.sact) {
            for (int col = rank; col < g.so4; col += tsz) {
                float4 ds = T[(d.gs_in + col) * rp + r];
                const float4 so = T[(c.s_out + col) * rp + r];
                ds.x *= act_bwd(g.sact, so.x); ds.y *= act_bwd(g.sact, so.y);
                ds.z *= act_bwd(g.sact, so.z); ds.w *= act_bwd(g.sact, so.w);
                T[(d.gs_in + col) * rp + r] = ds;
            }
        }
    }
    __syncthreads();
    if (work) {   // [dS_in ; dvn] = ds' . ws
        const int np_ = g.ksd4 * 4;
        for_blocks_team(g.ksd4, rank, tsz, [&](auto obc, int o0) {
            constexpr int OB = decltype(obc)::value;
            float2 acc[1][OB / 2];
            zero_acc(acc);
            rl_block<OB, 1>(T + d.gs_in * rp + r, 0, rp, g.so4, W + g.o_ws_b, np_, o0, acc);
#pragma unroll
            for (int j = 0; j < OB / 4; ++j)
                T[(d.gs_out + (o0 >> 2) + j) * rp + r] =
                    make_float4(acc[0][2 * j].x, acc[0][2 * j].y, acc[0][2 * j + 1].x, acc[0][2 * j + 1].y);
        });
    }
    __syncthreads();
    if (g.vi > 0) {
        if (work) {   // dVh = wv^T dVo + Vh * dvn / vn   (clamp passes the gradient where |Vh|^2 >= eps)
            const int hp = g.h4 * 4;
            for_blocks_team(g.h4, rank, tsz, [&](auto obc, int o0) {
                constexpr int OB = decltype(obc)::value;
                float2 acc[3][OB / 2];
                zero_acc(acc);
                if (g.vo > 0) rl_block<OB, 3>(T + d.gv_in * rp + r, d.gv_in_pc * rp, rp, g.vo4, W + g.o_wv_b, hp, o0, acc);
#pragma unroll
                for (int j = 0; j < OB / 4; ++j) {
                    const int col = (o0 >> 2) + j;
                    float4 vh[3], o[3];
#pragma unroll
                    for (int p = 0; p < 3; ++p) vh[p] = T[(c.vh + p * c.vh_pc + col) * rp + r];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int k = o0 + 4 * j + i;
                        const float x = f4get(vh[0], i), y = f4get(vh[1], i), z = f4get(vh[2], i);
                        const float q = x * x + y * y + z * z;
                        float f = 0.f;
                        if (k < g.h && q >= CGVP_EPS)
                            f = tile_at(T, rp, r, d.gs_out, g.si + k) / tile_at(T, rp, r, c.s_in, g.si + k);
                        const float2 a0 = acc[0][2 * j + (i >> 1)], a1 = acc[1][2 * j + (i >> 1)], a2 = acc[2][2 * j + (i >> 1)];
                        reinterpret_cast<float*>(&o[0])[i] = ((i & 1) ? a0.y : a0.x) + x * f;
                        reinterpret_cast<float*>(&o[1])[i] = ((i & 1) ? a1.y : a1.x) + y * f;
                        reinterpret_cast<float*>(&o[2])[i] = ((i & 1) ? a2.y : a2.x) + z * f;
                    }
#pragma unroll
                    for (int p = 0; p < 3; ++p) T[(d.dvh + p * d.dvh_pc + col) * rp + r] = o[p];
                }
            });
        }
        __syncthreads();
        if (work) {   // dV_in = wh^T dVh
            const int vip = g.vi4 * 4;
            for_blocks_team(g.vi4, rank, tsz, [&](auto obc, int o0) {
                constexpr int OB = decltype(obc)::value;
                float2 acc[3][OB / 2];
                zero_acc(acc);
                rl_block<OB, 3>(T + d.dvh * rp + r, d.dvh_pc * rp, rp, g.h4, W + g.o_wh_b, vip, o0, acc);
#pragma unroll
                for (int j = 0; j < OB / 4; ++j)
#pragma unroll
                    for (int p = 0; p < 3; ++p)
                        T[(d.gv_out + p * d.gv_out_pc + (o0 >> 2) + j) * rp + r] =
                            make_float4(acc[p][2 * j].x, acc[p][2 * j].y, acc[p][2 * j + 1].x, acc[p][2 * j + 1].y);
            });
        }
        __syncthreads();
    }
}

// ---- LayerNorm for one row (gvp_layers.py:231-242) ------------------------------------------------------------
// src -> dst (may alias).  Writes (mean, rstd, vector rms) to `stat` when stat >= 0.
__device__ __forceinline__ void ln_fwd_row(float4* T, int rp, int r, int ns, int nv, int s_src, int v_src, int vpc,
                                           int s_dst, int v_dst, const float* __restrict__ w,
                                           const float* __restrict__ b, int stat) {
    float mean = 0.f;
    for (int k = 0; k < ns; ++k) mean += tile_at(T, rp, r, s_src, k);
    mean /= (float)ns;
    float var = 0.f;
    for (int k = 0; k < ns; ++k) { const float d = tile_at(T, rp, r, s_src, k) - mean; var += d * d; }
    const float rstd = rsqrtf(var / (float)ns + CGVP_LN_EPS);
    for (int k = 0; k < ns; ++k)
        tile_at(T, rp, r, s_dst, k) = (tile_at(T, rp, r, s_src, k) - mean) * rstd * __ldg(w + k) + __ldg(b + k);
    float rms = 1.f;
    if (nv > 0) {
        float m = 0.f;
        for (int c = 0; c < nv; ++c) {
            const float x = tile_at(T, rp, r, v_src, c), y = tile_at(T, rp, r, v_src + vpc, c),
                        z = tile_at(T, rp, r, v_src + 2 * vpc, c);
            m += fmaxf(x * x + y * y + z * z, CGVP_EPS);                                                   // :240
        }
        rms = sqrtf(m / (float)nv);                                                                        // :241
        for (int p = 0; p < 3; ++p)
            for (int c = 0; c < nv; ++c) tile_at(T, rp, r, v_dst + p * vpc, c) = tile_at(T, rp, r, v_src + p * vpc, c) / rms;
    }
    if (stat >= 0) T[stat * rp + r] = make_float4(mean, rstd, rms, 0.f);
}

// Backward of the above.  x (pre-norm input) at (s_x, v_x); dy at (s_dy, v_dy); writes dx to (s_dx, v_dx)
// (may alias dy).  Parameter gradients are reduced across rows elsewhere (they need x_hat = (x-mean)*rstd).
__device__ __forceinline__ void ln_bwd_row(float4* T, int rp, int r, int ns, int nv, int s_x, int v_x, int vpc_x,
                                           int s_dy, int v_dy, int vpc_dy, int s_dx, int v_dx, int vpc_dx,
                                           const float* __restrict__ w, int stat) {
    const float4 st = T[stat * rp + r];
    const float mean = st.x, rstd = st.y, rms = st.z;
    float m1 = 0.f, m2 = 0.f;
    for (int k = 0; k < ns; ++k) {
        const float dyh = tile_at(T, rp, r, s_dy, k) * __ldg(w + k);
        const float xh = (tile_at(T, rp, r, s_x, k) - mean) * rstd;
        m1 += dyh; m2 += dyh * xh;
    }
    m1 /= (float)ns; m2 /= (float)ns;
    for (int k = 0; k < ns; ++k) {
        const float dyh = tile_at(T, rp, r, s_dy, k) * __ldg(w + k);
        const float xh = (tile_at(T, rp, r, s_x, k) - mean) * rstd;
        tile_at(T, rp, r, s_dx, k) = rstd * (dyh - m1 - xh * m2);
    }
    if (nv > 0) {
        float dot = 0.f;
        for (int p = 0; p < 3; ++p)
            for (int c = 0; c < nv; ++c) dot += tile_at(T, rp, r, v_dy + p * vpc_dy, c) * tile_at(T, rp, r, v_x + p * vpc_x, c);
        const float coef = dot / ((float)nv * rms * rms * rms);
        for (int c = 0; c < nv; ++c) {
            const float x = tile_at(T, rp, r, v_x, c), y = tile_at(T, rp, r, v_x + vpc_x, c),
                        z = tile_at(T, rp, r, v_x + 2 * vpc_x, c);
            const float pass = (x * x + y * y + z * z) >= CGVP_EPS ? coef : 0.f;
            tile_at(T, rp, r, v_dx, c) = tile_at(T, rp, r, v_dy, c) / rms - x * pass;
            tile_at(T, rp, r, v_dx + vpc_dx, c) = tile_at(T, rp, r, v_dy + vpc_dy, c) / rms - y * pass;
            tile_at(T, rp, r, v_dx + 2 * vpc_dx, c) = tile_at(T, rp, r, v_dy + 2 * vpc_dy, c) / rms - z * pass;
        }
    }
}

// ---- weight-gradient tile GEMM ----------------------------------------------------------------------------------
// G[a][b] += sum_{r < rows, p < np} A_p[r][a] * B_p[r][b]    (A = forward operand, K index; B = gradient, N index)
// expressed on 4x4 register blocks; blocks of ALL matrices of a chain are flattened into one list so that the
// CTA's threads share them evenly.  Result layout = the forward packed layout (G[a * ldg + b]).
struct DwMat {
    int col_a, pitch_a, na4;   // A columns (float4), plane pitch, count
    int col_b, pitch_b, nb4;
    int np;                    // planes (1 scalar, 3 vector)
    int goff;                  // float offset of this matrix in the chain's gradient arena
    int blk0;                  // first flattened block id
};
#define CGVP_MAX_DWMAT (4 * CGVP_MAX_CHAIN)
struct DwPlan {
    int nmat, nblk;
    DwMat m[CGVP_MAX_DWMAT];
};

constexpr CGVP_HD inline void dw_add(DwPlan& p, int col_a, int pitch_a, int na4, int col_b, int pitch_b, int nb4, int np, int goff) {
    if (na4 <= 0 || nb4 <= 0) return;
    DwMat& m = p.m[p.nmat++];
    m.col_a = col_a; m.pitch_a = pitch_a; m.na4 = na4; m.col_b = col_b; m.pitch_b = pitch_b; m.nb4 = nb4;
    m.np = np; m.goff = goff; m.blk0 = p.nblk;
    p.nblk += na4 * nb4;
}

// forward operands (A) and gradient operands (B) of GVP `g` with stage columns c / gradient columns d
constexpr CGVP_HD inline void dw_add_gvp(DwPlan& p, const GvpP& g, const StageIO& c, const GradIO& d, int goff) {
    if (g.vi > 0) dw_add(p, c.v_in, c.v_in_pc, g.vi4, d.dvh, d.dvh_pc, g.h4, 3, goff + g.o_wh_t);
    dw_add(p, c.s_in, 0, g.ks4, d.gs_in, 0, g.so4, 1, goff + g.o_ws_t);
    if (g.vi > 0 && g.vo > 0) dw_add(p, c.vh, c.vh_pc, g.h4, d.gv_in, d.gv_in_pc, g.vo4, 3, goff + g.o_wv_t);
    if (g.gate) dw_add(p, c.sp, 0, g.ksv4, d.dg, 0, g.vo4, 1, goff + g.o_wsv_t);
}

__device__ __forceinline__ void dw_block(const DwMat& m, int q, const float4* T, int rp, int rows, float (&acc)[16]) {
    const int a4 = q / m.nb4, b4 = q - a4 * m.nb4;
    for (int p = 0; p < m.np; ++p) {
        const float4* A = T + (m.col_a + p * m.pitch_a + a4) * rp;
        const float4* B = T + (m.col_b + p * m.pitch_b + b4) * rp;
        float2 c2[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) c2[i] = make_float2(acc[2 * i], acc[2 * i + 1]);
#pragma unroll 4
        for (int r = 0; r < rows; ++r) {
            const float4 a = A[r], b = B[r];
            fma2(c2[0], a.x, make_float2(b.x, b.y)); fma2(c2[1], a.x, make_float2(b.z, b.w));
            fma2(c2[2], a.y, make_float2(b.x, b.y)); fma2(c2[3], a.y, make_float2(b.z, b.w));
            fma2(c2[4], a.z, make_float2(b.x, b.y)); fma2(c2[5], a.z, make_float2(b.z, b.w));
            fma2(c2[6], a.w, make_float2(b.x, b.y)); fma2(c2[7], a.w, make_float2(b.z, b.w));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { acc[2 * i] = c2[i].x; acc[2 * i + 1] = c2[i].y; }
    }
}

// NSLOT register-resident blocks per thread (accumulated across all tiles of a persistent CTA); if the chain has
// more blocks than NSLOT * blockDim the remainder accumulates straight into this CTA's private global partial.
template <int NSLOT>
struct DwAcc {
    float acc[NSLOT][16];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int s = 0; s < NSLOT; ++s)
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[s][i] = 0.f;
    }
    __device__ __forceinline__ static const DwMat& find(const DwPlan& p, int q) {
        int m = 0;
        while (m + 1 < p.nmat && q >= p.m[m + 1].blk0) ++m;
        return p.m[m];
    }
    // one tile's contribution; `partial` = this CTA's arena (zero-initialised)
    __device__ __forceinline__ void tile(const DwPlan& p, const float4* T, int rp, int rows, float* partial) {
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
            const int q = threadIdx.x + s * blockDim.x;
            if (q < p.nblk) {
                const DwMat& m = find(p, q);
                dw_block(m, q - m.blk0, T, rp, rows, acc[s]);
            }
        }
        for (int q = threadIdx.x + NSLOT * blockDim.x; q < p.nblk; q += blockDim.x) {
            const DwMat& m = find(p, q);
            float a[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = 0.f;
            const int qq = q - m.blk0;
            dw_block(m, qq, T, rp, rows, a);
            const int a4 = qq / m.nb4, b4 = qq - a4 * m.nb4;
            float* gdst = partial + m.goff + (a4 * 4) * (m.nb4 * 4) + b4 * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4* row = reinterpret_cast<float4*>(gdst + i * m.nb4 * 4);
                float4 v = *row;
                v.x += a[4 * i]; v.y += a[4 * i + 1]; v.z += a[4 * i + 2]; v.w += a[4 * i + 3];
                *row = v;
            }
        }
    }
    __device__ __forceinline__ void flush(const DwPlan& p, float* partial) {
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
            const int q = threadIdx.x + s * blockDim.x;
            if (q < p.nblk) {
                const DwMat& m = find(p, q);
                const int qq = q - m.blk0;
                const int a4 = qq / m.nb4, b4 = qq - a4 * m.nb4;
                float* gdst = partial + m.goff + (a4 * 4) * (m.nb4 * 4) + b4 * 4;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    *reinterpret_cast<float4*>(gdst + i * m.nb4 * 4) =
                        make_float4(acc[s][4 * i], acc[s][4 * i + 1], acc[s][4 * i + 2], acc[s][4 * i + 3]);
            }
        }
    }
};

// ---- cooperative staging helpers --------------------------------------------------------------------------------
// copy the packed weight blocks of a chain into shared memory (float4 granularity)
__device__ __forceinline__ void copy_f4(float* dst, const float* __restrict__ src, int nfloats) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = threadIdx.x; i < (nfloats >> 2); i += blockDim.x) d4[i] = __ldg(s4 + i);
}

}  // namespace cgvp

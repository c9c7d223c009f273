// Row LayerNorm (nn.LayerNorm over the last dimension, eps inside the square root, affine) for the packed [rows, D] activations of
// the cross-attention block (preattn_norm / ff_norm of CrossAttentionModule, models/joint_gnn.py:321-408), forward and backward.
// One warp per row, lane = D / 32 consecutive floats (one coalesced load per row), mean / variance by shuffles.  The backward
// accumulates the gamma / beta gradients per lane in registers over the rows a warp walks, reduces them per CTA in shared
// memory and leaves one partial per CTA for a fixed-order final sum: deterministic, no atomics (the stock gamma/beta backward
// kernel alone takes ~60 us per call at 2 x 10^4 rows).
#include "cgvp_common.cuh"

template <int VEC>
__device__ __forceinline__ void ln_ld(const float* __restrict__ p, float (&x)[VEC]) {
    if constexpr (VEC % 4 == 0) {
#pragma unroll
        for (int i = 0; i < VEC / 4; ++i) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
            x[4 * i] = t.x; x[4 * i + 1] = t.y; x[4 * i + 2] = t.z; x[4 * i + 3] = t.w;
        }
    } else if constexpr (VEC == 2) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(p));
        x[0] = t.x; x[1] = t.y;
    } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) x[i] = __ldg(p + i);
    }
}
template <int VEC>
__device__ __forceinline__ void ln_st(float* __restrict__ p, const float (&x)[VEC]) {
    if constexpr (VEC % 4 == 0) {
#pragma unroll
        for (int i = 0; i < VEC / 4; ++i) reinterpret_cast<float4*>(p)[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
    } else if constexpr (VEC == 2) {
        *reinterpret_cast<float2*>(p) = make_float2(x[0], x[1]);
    } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) p[i] = x[i];
    }
}
__device__ __forceinline__ float ln_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int VEC>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, long long rows, float eps,
                                                      float* __restrict__ y, float* __restrict__ stats) {
    constexpr int D = 32 * VEC;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    float g[VEC], b[VEC];
    ln_ld<VEC>(gamma + lane * VEC, g);
    ln_ld<VEC>(beta + lane * VEC, b);
    for (long long r = warp; r < rows; r += nwarps) {
        float v[VEC];
        ln_ld<VEC>(x + r * D + lane * VEC, v);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) s += v[i];
        const float mean = ln_warp_sum(s) * (1.f / (float)D);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
        const float rstd = rsqrtf(ln_warp_sum(q) * (1.f / (float)D) + eps);
#pragma unroll
        for (int i = 0; i < VEC; ++i) v[i] = (v[i] - mean) * rstd * g[i] + b[i];
        ln_st<VEC>(y + r * D + lane * VEC, v);
        if (lane == 0) reinterpret_cast<float2*>(stats)[r] = make_float2(mean, rstd);
    }
}

template <int VEC>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                      const float* __restrict__ stats, const float* __restrict__ gamma, long long rows,
                                                      float* __restrict__ dx, float* __restrict__ partial) {
    constexpr int D = 32 * VEC;
    __shared__ float sm[8][2 * D];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    float g[VEC], dg[VEC], db[VEC];
    ln_ld<VEC>(gamma + lane * VEC, g);
#pragma unroll
    for (int i = 0; i < VEC; ++i) { dg[i] = 0.f; db[i] = 0.f; }
    for (long long r = warp; r < rows; r += nwarps) {
        float v[VEC], d[VEC];
        ln_ld<VEC>(x + r * D + lane * VEC, v);
        ln_ld<VEC>(dy + r * D + lane * VEC, d);
        const float2 st = __ldg(reinterpret_cast<const float2*>(stats) + r);
        float m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            v[i] = (v[i] - st.x) * st.y;                           // xhat
            dg[i] = fmaf(d[i], v[i], dg[i]);
            db[i] += d[i];
            d[i] *= g[i];
            m1 += d[i];
            m2 = fmaf(d[i], v[i], m2);
        }
        m1 = ln_warp_sum(m1) * (1.f / (float)D);
        m2 = ln_warp_sum(m2) * (1.f / (float)D);
#pragma unroll
        for (int i = 0; i < VEC; ++i) d[i] = st.y * (d[i] - m1 - v[i] * m2);
        ln_st<VEC>(dx + r * D + lane * VEC, d);
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) { sm[w][lane * VEC + i] = dg[i]; sm[w][D + lane * VEC + i] = db[i]; }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int ww = 0; ww < 8; ++ww) s += sm[ww][i];
        partial[(long long)blockIdx.x * 2 * D + i] = s;
    }
}

// dgamma / dbeta = sum of the per-CTA partials.  Block = 8 warps x 32 columns: warp w sums partials w, w + 8, ... of its
// lane's column (coalesced, 8 loads in flight), the eight slices meet in shared memory in warp order: deterministic.
__global__ void __launch_bounds__(256) ln_reduce_kernel(const float* __restrict__ partial, int parts, int D, float* __restrict__ dgamma,
                                                         float* __restrict__ dbeta) {
    __shared__ float sm[8][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;                       // column of the [parts][2D] partial matrix
    float s = 0.f;
    if (i < 2 * D) {
#pragma unroll 8
        for (int p = w; p < parts; p += 8) s += __ldg(partial + (long long)p * 2 * D + i);
    }
    sm[w][lane] = s;
    __syncthreads();
    if (w == 0 && i < 2 * D) {
        float t = 0.f;
#pragma unroll
        for (int ww = 0; ww < 8; ++ww) t += sm[ww][lane];
        (i < D ? dgamma[i] : dbeta[i - D]) = t;
    }
}

static int ln_grid(int64_t rows) {
    const long long want = (rows + 7) / 8, cap = (long long)cgvp_num_sms() * 4;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

extern "C" int32_t cgvp_layernorm_supported(int32_t D) { return (D == 32 || D == 64 || D == 128 || D == 256) ? 1 : 0; }

extern "C" int64_t cgvp_layernorm_workspace_bytes(int64_t rows, int32_t D) {
    if (!cgvp_layernorm_supported(D) || rows < 0) return -1;
    return (int64_t)ln_grid(rows) * 2 * D * 4 + 256;
}

static bool ln_al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int32_t cgvp_layernorm_fwd(const float* x, const float* gamma, const float* beta, int64_t rows, int32_t D, float eps,
                                      float* y, float* stats, void* stream) {
    CGVP_REQUIRE(cgvp_layernorm_supported(D) && rows >= 0, "layernorm_fwd: unsupported width %d", D);
    if (rows == 0) return 0;
    CGVP_REQUIRE(x && gamma && beta && y && stats, "layernorm_fwd: null argument");
    CGVP_REQUIRE(ln_al16(x) && ln_al16(gamma) && ln_al16(beta) && ln_al16(y) && ln_al16(stats), "layernorm_fwd: buffers must be 16-byte aligned");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = ln_grid(rows);
    if (D == 32) ln_fwd_kernel<1><<<grid, 256, 0, st>>>(x, gamma, beta, rows, eps, y, stats);
    else if (D == 64) ln_fwd_kernel<2><<<grid, 256, 0, st>>>(x, gamma, beta, rows, eps, y, stats);
    else if (D == 128) ln_fwd_kernel<4><<<grid, 256, 0, st>>>(x, gamma, beta, rows, eps, y, stats);
    else ln_fwd_kernel<8><<<grid, 256, 0, st>>>(x, gamma, beta, rows, eps, y, stats);
    CGVP_LAUNCH_CHECK("ln_fwd_kernel");
    return 0;
}

extern "C" int32_t cgvp_layernorm_bwd(const float* dy, const float* x, const float* stats, const float* gamma, int64_t rows, int32_t D,
                                      float* dx, float* dgamma, float* dbeta, void* ws, int64_t ws_bytes, void* stream) {
    CGVP_REQUIRE(cgvp_layernorm_supported(D) && rows >= 0, "layernorm_bwd: unsupported width %d", D);
    CGVP_REQUIRE(dy && x && stats && gamma && dx && dgamma && dbeta && ws, "layernorm_bwd: null argument");
    CGVP_REQUIRE(ln_al16(dy) && ln_al16(x) && ln_al16(stats) && ln_al16(gamma) && ln_al16(dx), "layernorm_bwd: buffers must be 16-byte aligned");
    CGVP_REQUIRE(ws_bytes >= cgvp_layernorm_workspace_bytes(rows, D), "layernorm_bwd: workspace too small");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    float* partial = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
    const int grid = ln_grid(rows);
    if (D == 32) ln_bwd_kernel<1><<<grid, 256, 0, st>>>(dy, x, stats, gamma, rows, dx, partial);
    else if (D == 64) ln_bwd_kernel<2><<<grid, 256, 0, st>>>(dy, x, stats, gamma, rows, dx, partial);
    else if (D == 128) ln_bwd_kernel<4><<<grid, 256, 0, st>>>(dy, x, stats, gamma, rows, dx, partial);
    else ln_bwd_kernel<8><<<grid, 256, 0, st>>>(dy, x, stats, gamma, rows, dx, partial);
    CGVP_LAUNCH_CHECK("ln_bwd_kernel");
    ln_reduce_kernel<<<cdiv(2 * D, 32), 256, 0, st>>>(partial, grid, D, dgamma, dbeta);
    CGVP_LAUNCH_CHECK("ln_reduce_kernel");
    return 0;
}

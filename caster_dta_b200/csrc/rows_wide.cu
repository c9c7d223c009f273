// Wide node update (forward): the GVPConvLayer tail  x <- LN1(x1 + D1(FF(x1))),  x1 = LN0(x + D0(dh))
// (models/gvp_layers.py:407-410, ff_func :355-364) for node dims too wide for a register-resident row
// (e.g. BASELINE config 5: nodes (100,16), feed-forward hidden (400,32)).  fp32 throughout.
//
// The two feed-forward GVPs are dominated by their scalar projections W_s (133 x 400 and 433 x 100 at config-5 dims),
// which are plain dense GEMMs over all N nodes; they run in a register-tiled FFMA GEMM.  Everything else is per-node
// vector work (LayerNorm, W_h / W_mu on 3 x 16..32 values, norms, gates) in three warp-per-node kernels:
//     wide_k1 : x1 = LN0(x + m0 * dh);  Vh0 = W_h0 V1;  A0 = [s1 ; |Vh0|]
//     GEMM    : [relu(s'0) | gate0] = A0 [W_s0 ; W_sv0 W_s0]^T + b          (gate folded: it reads the PRE-activation s')
//     wide_k2 : V_mid = (W_mu0 Vh0) * sigmoid(gate0);  Vh1 = W_h1 V_mid;  A1 = [relu(s'0) ; |Vh1|]   (in place)
//     GEMM    : [s'1 | gate1] = A1 [W_s1 ; W_sv1 W_s1]^T + b
//     wide_k3 : ff = (s'1, (W_mu1 Vh1) * sigmoid(gate1));  out = LN1(x1 + m1 * ff)
// Dims are runtime values: the path serves any (ns, nv) node-update descriptor with gated, vector_act = None GVPs.
#include "cgvp_common.cuh"

struct WideDims {
    int ns, nv;          // node dims
    int hs, hv;          // feed-forward hidden dims (so, vo of GVP 0)
    int h0, h1;          // hidden vector channels of the two GVPs
    int sact0;           // scalar activation of GVP 0 (GVP 1 has none)
    // packed fp32 block offsets of the two GVPs (cgvp_common.cuh)
    GvpP g0, g1;
    int ka0, ka1;        // GEMM K: ns + h0, hs + h1
    int n0, n1;          // GEMM N: hs + hv, ns + nv
};

// ---- register-tiled FFMA GEMM: C[M, N] = act(A[M, K] B[K, N] + bias), relu on columns < relu_cols -----------------------
// 128 x 64 tile per CTA, 8 x 4 outputs per thread, K in steps of 16; the next K tile is fetched into registers
// (float4 loads where lda / ldb / K allow) while the current one is multiplied out of shared memory.
#define WG_BM 128
#define WG_BN 64
#define WG_BK 16
__global__ void __launch_bounds__(256) wide_gemm_kernel(long long M, int N, int K, const float* __restrict__ A, int lda,
                                                        const float* __restrict__ B, int ldb, const float* __restrict__ bias,
                                                        float* __restrict__ C, int ldc, int relu_cols) {
    __shared__ __align__(16) float As[WG_BK][WG_BM + 4];
    __shared__ __align__(16) float Bs[WG_BK][WG_BN];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;       // 16 x 16 threads
    const long long m0 = (long long)blockIdx.x * WG_BM;
    const int n0 = blockIdx.y * WG_BN;
    const bool a_vec = (lda & 3) == 0 && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    const bool b_vec = (ldb & 3) == 0 && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
    // this thread's pieces of a K tile: A rows ar0, ar0 + 64 at k offset ak (4 values each); B row bk, 4 columns at bn
    const int ar0 = tid >> 2, ak = (tid & 3) * 4;
    const int bk = tid >> 4, bn = (tid & 15) * 4;
    float4 ra[2], rb;
    auto fetch = [&](int k0) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const long long m = m0 + ar0 + 64 * q;
            const int k = k0 + ak;
            if (m < M && a_vec && k + 3 < K) ra[q] = __ldg(reinterpret_cast<const float4*>(A + m * lda + k));
            else {
                float t[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) t[j] = (m < M && k + j < K) ? __ldg(A + m * lda + k + j) : 0.f;
                ra[q] = make_float4(t[0], t[1], t[2], t[3]);
            }
        }
        const int k = k0 + bk, n = n0 + bn;
        if (k < K && b_vec && n + 3 < N) rb = __ldg(reinterpret_cast<const float4*>(B + (long long)k * ldb + n));
        else {
            float t[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) t[j] = (k < K && n + j < N) ? __ldg(B + (long long)k * ldb + n + j) : 0.f;
            rb = make_float4(t[0], t[1], t[2], t[3]);
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int r = ar0 + 64 * q;
            As[ak][r] = ra[q].x; As[ak + 1][r] = ra[q].y; As[ak + 2][r] = ra[q].z; As[ak + 3][r] = ra[q].w;
        }
        *reinterpret_cast<float4*>(&Bs[bk][bn]) = rb;
    };
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    fetch(0);
    for (int k0 = 0; k0 < K; k0 += WG_BK) {
        stash();
        __syncthreads();
        if (k0 + WG_BK < K) fetch(k0 + WG_BK);               // in flight while this tile is multiplied
#pragma unroll
        for (int k = 0; k < WG_BK; ++k) {
            // rows ty*4 .. +3 and 64 + ty*4 .. +3: conflict-free float4 reads
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    const bool c_vec = (ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= M) continue;
        const int n = n0 + tx * 4;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[j] = acc[i][j] + ((bias && n + j < N) ? __ldg(bias + n + j) : 0.f);
            if (n + j < relu_cols) v[j] = fmaxf(v[j], 0.f);
        }
        if (c_vec && n + 3 < N) *reinterpret_cast<float4*>(C + m * ldc + n) = make_float4(v[0], v[1], v[2], v[3]);
        else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n + j < N) C[m * ldc + n + j] = v[j];
        }
    }
}

// ---- GEMM weights: [W_s ; W_sv W_s] (K-major) and the fused bias ------------------------------------------------------------
// Wf[k][n], k < ksd (rows of ws_t), n < so + vo;  bf[n].   n < so: ws_t[k][n] / bias;  else the gate column o = n - so:
//   sum_j ws_t[k][j] wsv_t[j][o]   /   sum_j bias[j] wsv_t[j][o] + gate bias[o]
__global__ void wide_pack_kernel(const float* __restrict__ w, GvpP g, float* __restrict__ Wf, float* __restrict__ bf) {
    const int ksd = g.si + g.h, n_out = g.so + g.vo, sop = g.so4 * 4, vop = g.vo4 * 4;
    const int total = (ksd + 1) * n_out;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int k = i / n_out, n = i % n_out;           // k == ksd: the bias row
        float v;
        if (n < g.so) v = w[g.o_ws_t + k * sop + n];
        else {
            const int o = n - g.so;
            v = k == ksd ? w[g.o_wsv_t + g.so * vop + o] : 0.f;
            for (int j = 0; j < g.so; ++j) v += w[g.o_ws_t + k * sop + j] * w[g.o_wsv_t + j * vop + o];
        }
        if (k < ksd) Wf[(long long)k * n_out + n] = v;
        else bf[n] = v;
    }
}

// ---- warp-per-node kernels ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// LayerNorm of one node held in shared memory by its warp: s[0..S) in place; V[c][p] (c < C) divided by the channel rms
__device__ __forceinline__ void warp_layer_norm(float* s, float* v, int S, int C, const float* __restrict__ w,
                                                const float* __restrict__ b, int lane) {
    float sum = 0.f;
    for (int k = lane; k < S; k += 32) sum += s[k];
    const float mean = warp_sum(sum) / (float)S;
    float var = 0.f;
    for (int k = lane; k < S; k += 32) { const float d = s[k] - mean; var += d * d; }
    const float rstd = rsqrtf(warp_sum(var) / (float)S + CGVP_LN_EPS);
    for (int k = lane; k < S; k += 32) s[k] = (s[k] - mean) * rstd * __ldg(w + k) + __ldg(b + k);
    if (C > 0) {
        float m = 0.f;
        for (int c = lane; c < C; c += 32) m += fmaxf(v[3 * c] * v[3 * c] + v[3 * c + 1] * v[3 * c + 1] + v[3 * c + 2] * v[3 * c + 2], CGVP_EPS);
        const float rms = sqrtf(warp_sum(m) / (float)C);
        for (int j = lane; j < 3 * C; j += 32) v[j] /= rms;
    }
}

struct WideArgs {
    long long N;
    WideDims d;
    const float *x_s, *x_v, *h_s, *h_v, *m0s, *m0v, *m1s, *m1v, *ln0_w, *ln0_b, *ln1_w, *ln1_b;
    const float *w0, *w1;          // generic packed fp32 blocks of the two GVPs
    float *x1s, *x1v;              // [N, ns], [N, nv, 3]
    float *a0;                     // [N, ka0]
    float *vh0;                    // [N, 3, h0]
    float *sg0;                    // [N, n0]  -> becomes A1 = [relu(s'0) ; vn1] in place (ka1 == n0)
    float *vh1;                    // [N, 3, h1]
    float *sg1;                    // [N, n1]
    float *out_s, *out_v;
};

#define WIDE_WARPS 8
// per-warp shared scratch (floats): scalars + vectors of one node
__global__ void __launch_bounds__(WIDE_WARPS * 32) wide_k1_kernel(const WideArgs a, int scratch) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* s = sm + warp * scratch;
    float* v = s + a.d.ns;
    const int ns = a.d.ns, nv = a.d.nv, h0 = a.d.h0;
    const float* wh = a.w0 + a.d.g0.o_wh_t;               // [vi_p][h_p], k-major
    const int hp = a.d.g0.h4 * 4;
    for (long long n = (long long)blockIdx.x * WIDE_WARPS + warp; n < a.N; n += (long long)gridDim.x * WIDE_WARPS) {
        for (int k = lane; k < ns; k += 32)
            s[k] = a.x_s[n * ns + k] + (a.m0s ? a.m0s[n * ns + k] : 1.f) * a.h_s[n * ns + k];
        for (int j = lane; j < 3 * nv; j += 32)
            v[j] = a.x_v[n * 3 * nv + j] + (a.m0v ? a.m0v[n * nv + j / 3] : 1.f) * a.h_v[n * 3 * nv + j];
        __syncwarp();
        warp_layer_norm(s, v, ns, nv, a.ln0_w, a.ln0_b, lane);
        __syncwarp();
        for (int k = lane; k < ns; k += 32) { a.x1s[n * ns + k] = s[k]; a.a0[n * a.d.ka0 + k] = s[k]; }
        for (int j = lane; j < 3 * nv; j += 32) a.x1v[n * 3 * nv + j] = v[j];
        for (int o = lane; o < h0; o += 32) {                                        // Vh0 = W_h0 V1, |Vh0|     :152-153
            float x = 0.f, y = 0.f, z = 0.f;
            for (int c = 0; c < nv; ++c) {
                const float w = __ldg(wh + c * hp + o);
                x = fmaf(v[3 * c], w, x); y = fmaf(v[3 * c + 1], w, y); z = fmaf(v[3 * c + 2], w, z);
            }
            a.vh0[(n * 3 + 0) * h0 + o] = x; a.vh0[(n * 3 + 1) * h0 + o] = y; a.vh0[(n * 3 + 2) * h0 + o] = z;
            a.a0[n * a.d.ka0 + ns + o] = sqrtf(fmaxf(x * x + y * y + z * z, CGVP_EPS));
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(WIDE_WARPS * 32) wide_k2_kernel(const WideArgs a, int scratch) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* vh = sm + warp * scratch;                      // [3][h0]
    float* vm = vh + 3 * a.d.h0;                          // [3][hv]  V_mid
    const int hs = a.d.hs, hv = a.d.hv, h0 = a.d.h0, h1 = a.d.h1;
    const float* wv = a.w0 + a.d.g0.o_wv_t;               // [h_p][vo_p]
    const int vop = a.d.g0.vo4 * 4;
    const float* wh1 = a.w1 + a.d.g1.o_wh_t;              // [vi_p][h_p]
    const int hp1 = a.d.g1.h4 * 4;
    for (long long n = (long long)blockIdx.x * WIDE_WARPS + warp; n < a.N; n += (long long)gridDim.x * WIDE_WARPS) {
        for (int j = lane; j < 3 * h0; j += 32) vh[j] = a.vh0[n * 3 * h0 + j];
        __syncwarp();
        for (int o = lane; o < hv; o += 32) {                                        // Vo0 = W_mu0 Vh0, gated      :156-163
            const float g = 1.f / (1.f + expf(-a.sg0[n * a.d.n0 + hs + o]));
            float x = 0.f, y = 0.f, z = 0.f;
            for (int k = 0; k < h0; ++k) {
                const float w = __ldg(wv + k * vop + o);
                x = fmaf(vh[k], w, x); y = fmaf(vh[h0 + k], w, y); z = fmaf(vh[2 * h0 + k], w, z);
            }
            vm[o] = x * g; vm[hv + o] = y * g; vm[2 * hv + o] = z * g;
        }
        __syncwarp();
        for (int o = lane; o < h1; o += 32) {                                        // Vh1 = W_h1 V_mid, |Vh1|
            float x = 0.f, y = 0.f, z = 0.f;
            for (int c = 0; c < hv; ++c) {
                const float w = __ldg(wh1 + c * hp1 + o);
                x = fmaf(vm[c], w, x); y = fmaf(vm[hv + c], w, y); z = fmaf(vm[2 * hv + c], w, z);
            }
            a.vh1[(n * 3 + 0) * h1 + o] = x; a.vh1[(n * 3 + 1) * h1 + o] = y; a.vh1[(n * 3 + 2) * h1 + o] = z;
            a.sg0[n * a.d.n0 + hs + o] = sqrtf(fmaxf(x * x + y * y + z * z, CGVP_EPS));   // overwrites the consumed gate column
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(WIDE_WARPS * 32) wide_k3_kernel(const WideArgs a, int scratch) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* s = sm + warp * scratch;                       // [ns]
    float* v = s + a.d.ns;                                // [nv][3]
    float* vh = v + 3 * a.d.nv;                           // [3][h1]
    const int ns = a.d.ns, nv = a.d.nv, h1 = a.d.h1;
    const float* wv = a.w1 + a.d.g1.o_wv_t;               // [h_p][vo_p]
    const int vop = a.d.g1.vo4 * 4;
    for (long long n = (long long)blockIdx.x * WIDE_WARPS + warp; n < a.N; n += (long long)gridDim.x * WIDE_WARPS) {
        for (int j = lane; j < 3 * h1; j += 32) vh[j] = a.vh1[n * 3 * h1 + j];
        for (int k = lane; k < ns; k += 32)
            s[k] = a.x1s[n * ns + k] + (a.m1s ? a.m1s[n * ns + k] : 1.f) * a.sg1[n * a.d.n1 + k];
        __syncwarp();
        for (int o = lane; o < nv; o += 32) {                                        // ff_V = (W_mu1 Vh1) * sigmoid(gate1)
            const float g = 1.f / (1.f + expf(-a.sg1[n * a.d.n1 + ns + o]));
            float x = 0.f, y = 0.f, z = 0.f;
            for (int k = 0; k < h1; ++k) {
                const float w = __ldg(wv + k * vop + o);
                x = fmaf(vh[k], w, x); y = fmaf(vh[h1 + k], w, y); z = fmaf(vh[2 * h1 + k], w, z);
            }
            const float m = a.m1v ? a.m1v[n * nv + o] : 1.f;
            v[3 * o] = a.x1v[(n * nv + o) * 3] + m * x * g;
            v[3 * o + 1] = a.x1v[(n * nv + o) * 3 + 1] + m * y * g;
            v[3 * o + 2] = a.x1v[(n * nv + o) * 3 + 2] + m * z * g;
        }
        __syncwarp();
        warp_layer_norm(s, v, ns, nv, a.ln1_w, a.ln1_b, lane);
        __syncwarp();
        for (int k = lane; k < ns; k += 32) a.out_s[n * ns + k] = s[k];
        for (int j = lane; j < 3 * nv; j += 32) a.out_v[n * 3 * nv + j] = v[j];
        __syncwarp();
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------------------
bool cgvp_fast_paths_enabled();
int64_t gemm_tc_packed_floats(int N, int K);
int gemm_tc_pack(const float* w, long long sn, long long sk, int N, int K, float* whi, float* wlo, cudaStream_t st);
int gemm_tc(long long M, int N, int K, const float* A, long long lda, const float* whi, const float* wlo, const float* bias,
            float* C, long long ldc, int relu_cols, cudaStream_t st);
static bool g_wide_tensor_gemm = true;      // fp32-accurate 3xTF32 tcgen05 GEMM (gemm_tc.cu) instead of the FFMA GEMM
extern "C" int32_t cgvp_set_wide_gemm(int32_t tensor) { g_wide_tensor_gemm = tensor != 0; return 0; }

static bool wide_matches(const CgvpRowDesc* d, WideDims& w) {
    if (!(d->has_residual_in && d->pre_norm && d->post_residual && d->post_norm && d->n_gvp == 2 && d->onehot == 0)) return false;
    const CgvpGvpDesc &a = d->gvp[0], &b = d->gvp[1];
    if (d->in_s < 64 || d->in_v <= 0) return false;                      // narrow rows: the register / tile kernels are better
    if (a.si != d->in_s || a.vi != d->in_v || b.si != a.so || b.vi != a.vo || b.so != d->in_s || b.vo != d->in_v) return false;
    if (!a.vector_gate || !b.vector_gate || a.vector_act != CGVP_ACT_NONE || b.vector_act != CGVP_ACT_NONE) return false;
    if (b.scalar_act != CGVP_ACT_NONE || (a.scalar_act != CGVP_ACT_RELU && a.scalar_act != CGVP_ACT_NONE)) return false;
    if (a.vo <= 0) return false;
    w.ns = d->in_s; w.nv = d->in_v; w.hs = a.so; w.hv = a.vo; w.h0 = a.h; w.h1 = b.h; w.sact0 = a.scalar_act;
    w.g0 = make_gvp_p(a); w.g1 = make_gvp_p(b);
    w.ka0 = w.ns + w.h0; w.n0 = w.hs + w.hv; w.ka1 = w.hs + w.h1; w.n1 = w.ns + w.nv;
    return w.ka1 == w.n0;                                                // A1 reuses the [s' | gate] buffer in place (h1 == hv)
}

static int64_t wide_layout(const WideDims& w, int64_t N, int64_t* off /*[12]*/) {
    int64_t o = 0;
    auto take = [&](int64_t floats) { const int64_t r = o; o += align_up(floats * 4, 256); return r; };
    off[0] = take((int64_t)w.ka0 * w.n0); off[1] = take(w.n0);          // Wf0, bf0
    off[2] = take((int64_t)w.ka1 * w.n1); off[3] = take(w.n1);          // Wf1, bf1
    off[4] = take(N * w.ns); off[5] = take(N * 3 * w.nv);               // x1s, x1v
    off[6] = take(N * w.ka0); off[7] = take(N * 3 * w.h0);              // a0, vh0
    off[8] = take(N * w.n0); off[9] = take(N * 3 * w.h1);               // sg0 / a1, vh1
    off[10] = take(N * w.n1);                                           // sg1
    off[11] = take(2 * (gemm_tc_packed_floats(w.n0, w.ka0) + gemm_tc_packed_floats(w.n1, w.ka1)));   // split weights (hi, lo) x 2 GEMMs
    return o + 256;
}

int64_t rows_wide_workspace_bytes(const CgvpRowDesc* desc, int64_t rows) {
    WideDims w;
    if (!wide_matches(desc, w)) return 0;
    int64_t off[12];
    return wide_layout(w, rows, off);
}

// Returns 1 if the wide path served the call (*rc = result), 0 if another kernel family must run.
int rows_fwd_wide(const CgvpRowDesc* desc, const CgvpRowArgs* args, void* ws, int64_t ws_bytes, cudaStream_t st, int* rc) {
    WideDims w;
    if (!cgvp_fast_paths_enabled() || args->rows <= 0 || args->in_index || !wide_matches(desc, w)) return 0;
    const int64_t N = args->rows;
    int64_t off[12];
    const int64_t need = wide_layout(w, N, off);
    if (!ws || ws_bytes < need) return 0;
    *rc = 0;
    char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
    auto F = [&](int i) { return reinterpret_cast<float*>(base + off[i]); };
    auto fail = [&](const char* what) {
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) return false;
        cgvp_set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
        *rc = (int)e;
        return true;
    };
    WideArgs a;
    memset(&a, 0, sizeof(a));
    a.N = N; a.d = w;
    a.x_s = args->in_s; a.x_v = args->in_v; a.h_s = args->h_s; a.h_v = args->h_v;
    a.m0s = args->mask0_s; a.m0v = args->mask0_v; a.m1s = args->mask1_s; a.m1v = args->mask1_v;
    a.ln0_w = args->ln0_w; a.ln0_b = args->ln0_b; a.ln1_w = args->ln1_w; a.ln1_b = args->ln1_b;
    a.w0 = args->h_packed[0]; a.w1 = args->h_packed[1];
    a.x1s = F(4); a.x1v = F(5); a.a0 = F(6); a.vh0 = F(7); a.sg0 = F(8); a.vh1 = F(9); a.sg1 = F(10);
    a.out_s = args->out_s; a.out_v = args->out_v;
    const int sms = cgvp_num_sms() > 0 ? cgvp_num_sms() : 148;
    wide_pack_kernel<<<64, 256, 0, st>>>(a.w0, w.g0, F(0), F(1));
    wide_pack_kernel<<<64, 256, 0, st>>>(a.w1, w.g1, F(2), F(3));
    if (fail("wide_pack_kernel")) return 1;
    const int grid = (int)(cdiv64(N, WIDE_WARPS) < (int64_t)sms * 8 ? cdiv64(N, WIDE_WARPS) : (int64_t)sms * 8);
    const int sc1 = w.ns + 3 * w.nv, sc2 = 3 * w.h0 + 3 * w.hv, sc3 = w.ns + 3 * w.nv + 3 * w.h1;
    const bool tensor = g_wide_tensor_gemm && (w.ka0 & 3) == 0 && (w.n0 & 3) == 0;
    float *whi0 = F(11), *wlo0 = whi0 + gemm_tc_packed_floats(w.n0, w.ka0), *whi1 = wlo0 + gemm_tc_packed_floats(w.n0, w.ka0),
          *wlo1 = whi1 + gemm_tc_packed_floats(w.n1, w.ka1);
    if (tensor) {                                          // Wf is [K][N]: W(n, k) = Wf[k * N + n]
        if ((*rc = gemm_tc_pack(F(0), 1, w.n0, w.n0, w.ka0, whi0, wlo0, st))) return 1;
        if ((*rc = gemm_tc_pack(F(2), 1, w.n1, w.n1, w.ka1, whi1, wlo1, st))) return 1;
    }
    const int relu0 = w.sact0 == CGVP_ACT_RELU ? w.hs : 0;
    cgvp_prof_begin(CGVP_K_ROWS_FWD, st);
    wide_k1_kernel<<<grid, WIDE_WARPS * 32, (size_t)WIDE_WARPS * sc1 * 4, st>>>(a, sc1);
    if (tensor) { if ((*rc = gemm_tc(N, w.n0, w.ka0, a.a0, w.ka0, whi0, wlo0, F(1), a.sg0, w.n0, relu0, st))) return 1; }
    else wide_gemm_kernel<<<dim3((unsigned)cdiv64(N, WG_BM), (unsigned)cdiv(w.n0, WG_BN)), 256, 0, st>>>(
            N, w.n0, w.ka0, a.a0, w.ka0, F(0), w.n0, F(1), a.sg0, w.n0, relu0);
    wide_k2_kernel<<<grid, WIDE_WARPS * 32, (size_t)WIDE_WARPS * sc2 * 4, st>>>(a, sc2);
    if (tensor) { if ((*rc = gemm_tc(N, w.n1, w.ka1, a.sg0, w.n0, whi1, wlo1, F(3), a.sg1, w.n1, 0, st))) return 1; }
    else wide_gemm_kernel<<<dim3((unsigned)cdiv64(N, WG_BM), (unsigned)cdiv(w.n1, WG_BN)), 256, 0, st>>>(
            N, w.n1, w.ka1, a.sg0, w.n0, F(2), w.n1, F(3), a.sg1, w.n1, 0);
    wide_k3_kernel<<<grid, WIDE_WARPS * 32, (size_t)WIDE_WARPS * sc3 * 4, st>>>(a, sc3);
    cgvp_prof_end(CGVP_K_ROWS_FWD, st);
    fail("wide node-update kernels");
    return 1;
}

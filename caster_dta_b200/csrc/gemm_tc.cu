// fp32-accurate dense GEMM on the tensor cores (tcgen05, kind::tf32, split precision "3xTF32"), sm_100a:
//     C[M, N] = act(A[M, K] . W[N, K]^T + bias)
// Each fp32 operand is split into hi = tf32(x) and lo = x - hi; the product is accumulated in fp32 (TMEM) as
// hi*hi + hi*lo + lo*hi, i.e. with ~2^-21 relative error per product instead of TF32's 2^-11 -- the result meets the
// fp32 parity bound (<= 1e-4, measured ~1e-6) while the work runs on the tensor pipe, which has ~10x the FFMA rate even
// at three MMAs per product.
// Used for the dense W_s projections of the wide node update (rows_wide.cu); the building block for moving the fused
// GVP chains to fp32-accurate tensor-core arithmetic.
//
// CTA = 128 threads, one 128 x NT output tile (NT <= 256), K in chunks of 32:  A chunk: coalesced global loads,
// split, stored in the K-major no-swizzle core-matrix layout [k/4][row][4];  W chunk: pre-split / pre-packed in the same
// layout by gemm_tc_pack_kernel and fetched with cp.async.bulk on an mbarrier;  12 tcgen05.mma per chunk;  epilogue on
// tcgen05.ld rows.  Two CTAs per SM overlap each other's load / MMA / epilogue phases.
#include "cgvp_common.cuh"
#include "cgvp_tc.cuh"

#define GT_KC 32                       // K per chunk (floats)
#define GT_LBO_A (2048 + 64)           // byte pitch between 16-byte k-chunks of the A tile (+64: conflict-free transposing stores)

struct GemmTcArgs {
    long long M;
    int N, K, Npad, Kpad, NT;          // Npad % 16 == 0, Kpad % GT_KC == 0, NT = N tile (<= 256, % 16 == 0)
    const float* A;
    long long lda;
    const float *whi, *wlo;            // packed [Kpad/4][Npad][4]
    const float* bias;
    float* C;
    long long ldc;
    int relu_cols;
};

using tcx::tf32_hi;
using tcx::idesc_tf32;
using tcx::mma_tf32;

// W(n, k) = w[n * sn + k * sk]  ->  hi / lo in [k/4][Npad][4], zero padded
__global__ void gemm_tc_pack_kernel(const float* __restrict__ w, long long sn, long long sk, int N, int K, int Npad, int Kpad,
                                    float* __restrict__ whi, float* __restrict__ wlo) {
    const long long total = (long long)Kpad * Npad;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int kk = (int)(i & 3);
        const long long j = i >> 2;
        const int n = (int)(j % Npad), k = (int)(j / Npad) * 4 + kk;
        const float x = (n < N && k < K) ? w[n * sn + k * sk] : 0.f;
        const float h = tf32_hi(x);
        whi[i] = h;
        wlo[i] = x - h;
    }
}

__global__ void __launch_bounds__(128, 2) gemm_tc_kernel(const __grid_constant__ GemmTcArgs a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int NT = a.NT;
    unsigned char* A_hi = smem;
    unsigned char* A_lo = A_hi + 8 * GT_LBO_A;
    unsigned char* B_hi = A_lo + 8 * GT_LBO_A;
    unsigned char* B_lo = B_hi + 8 * NT * 16;
    uint64_t* bbar = reinterpret_cast<uint64_t*>(B_lo + 8 * NT * 16);
    uint64_t* mbar = bbar + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(mbar + 1);
    const int tmem_cols = NT <= 32 ? 32 : (NT <= 64 ? 64 : (NT <= 128 ? 128 : 256));
    if (tid == 0) { tcx::mbar_init(bbar, 1); tcx::mbar_init(mbar, 1); tcx::fence_mbar_init(); }
    if (warp == 0) tcx::tmem_alloc(slot, tmem_cols);
    tcx::tc_fence_before();
    __syncthreads();
    tcx::tc_fence_after();
    const uint32_t tm0 = *slot;
    const uint32_t tm = tm0 + ((uint32_t)(warp * 32) << 16);
    const long long m0 = (long long)blockIdx.x * 128;
    const int n0 = blockIdx.y * NT;
    const uint32_t id = idesc_tf32(NT);
    const bool a_vec = (a.lda & 3) == 0 && (reinterpret_cast<uintptr_t>(a.A) & 15) == 0;
    uint32_t pb = 0, pm = 0;
    const int nchunks = a.Kpad / GT_KC;
    for (int c = 0; c < nchunks; ++c) {
        // ---- A chunk: 128 rows x 32 floats = 1024 float4, 8 per thread; lane -> (row, k4) with k4 fastest (coalesced)
        const int k0 = c * GT_KC;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int i = tid + q * 128, r = i >> 3, k4 = i & 7;
            const long long m = m0 + r;
            const int k = k0 + 4 * k4;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < a.M) {
                if (a_vec && k + 3 < a.K) x = __ldg(reinterpret_cast<const float4*>(a.A + m * a.lda + k));
                else {
                    if (k < a.K) x.x = __ldg(a.A + m * a.lda + k);
                    if (k + 1 < a.K) x.y = __ldg(a.A + m * a.lda + k + 1);
                    if (k + 2 < a.K) x.z = __ldg(a.A + m * a.lda + k + 2);
                    if (k + 3 < a.K) x.w = __ldg(a.A + m * a.lda + k + 3);
                }
            }
            const float4 h = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
            *reinterpret_cast<float4*>(A_hi + k4 * GT_LBO_A + r * 16) = h;
            *reinterpret_cast<float4*>(A_lo + k4 * GT_LBO_A + r * 16) = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
        }
        // ---- W chunk: 8 k4-slices of NT rows, hi and lo, through the TMA engine
        if (tid == 0) {
            tcx::mbar_expect_tx(bbar, (uint32_t)(16 * NT * 16));
            for (int k4 = 0; k4 < 8; ++k4) {
                const size_t off = ((size_t)(k0 / 4 + k4) * a.Npad + n0) * 4;
                tcx::bulk_g2s(B_hi + k4 * NT * 16, a.whi + off, (uint32_t)(NT * 16), bbar);
                tcx::bulk_g2s(B_lo + k4 * NT * 16, a.wlo + off, (uint32_t)(NT * 16), bbar);
            }
        }
        tcx::fence_proxy_async();
        tcx::tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tcx::mbar_wait(bbar, pb);
            tcx::tc_fence_after();
            const uint32_t ah = tcx::smem_u32(A_hi), al = tcx::smem_u32(A_lo), bh = tcx::smem_u32(B_hi), bl = tcx::smem_u32(B_lo);
#pragma unroll
            for (int k8 = 0; k8 < GT_KC / 8; ++k8) {
                const uint64_t dah = tcx::smem_desc(ah + k8 * 2 * GT_LBO_A, GT_LBO_A), dal = tcx::smem_desc(al + k8 * 2 * GT_LBO_A, GT_LBO_A);
                const uint64_t dbh = tcx::smem_desc(bh + k8 * 2 * NT * 16, NT * 16), dbl = tcx::smem_desc(bl + k8 * 2 * NT * 16, NT * 16);
                mma_tf32(tm0, dal, dbh, id, (c | k8) != 0);      // small terms first
                mma_tf32(tm0, dah, dbl, id, 1);
                mma_tf32(tm0, dah, dbh, id, 1);
            }
            tcx::mma_commit(mbar);
        }
        pb ^= 1;
        tcx::mbar_wait(mbar, pm);                               // the chunk's MMAs are done: the tiles may be overwritten
        pm ^= 1;
        tcx::tc_fence_after();
    }
    // ---- epilogue: this thread's row, 16 columns at a time
    const long long m = m0 + tid;
    for (int cb = 0; cb < NT; cb += 16) {
        float d[16];
        tcx::tmem_ld16(tm + cb, d);
        tcx::tmem_ld_wait(d);
        if (m < a.M) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                const int n = n0 + cb + 4 * j4;
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[j] = d[4 * j4 + j] + ((a.bias && n + j < a.N) ? __ldg(a.bias + n + j) : 0.f);
                    if (n + j < a.relu_cols) v[j] = fmaxf(v[j], 0.f);
                }
                if (n + 3 < a.N && (a.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(a.C) & 15) == 0)
                    *reinterpret_cast<float4*>(a.C + m * a.ldc + n) = make_float4(v[0], v[1], v[2], v[3]);
                else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (n + j < a.N) a.C[m * a.ldc + n + j] = v[j];
                }
            }
        }
    }
    tcx::tc_fence_before();
    __syncthreads();
    if (warp == 0) tcx::tmem_dealloc(tm0, tmem_cols);
}

// ---- host side -------------------------------------------------------------------------------------------------------------
// N is covered by equal tiles of NT columns (multiple of 16, <= 256); the packed weights are padded to NT * tiles rows.
static void gemm_tc_tiling(int N, int* NT, int* Npad) {
    const int n16 = (int)align_up(N, 16), tiles = cdiv(n16, 256);
    *NT = (int)align_up(cdiv(n16, tiles), 16);
    *Npad = *NT * tiles;
}
int64_t gemm_tc_packed_floats(int N, int K) {
    int NT, Npad;
    gemm_tc_tiling(N, &NT, &Npad);
    return (int64_t)align_up(K, GT_KC) * Npad;
}

// W(n, k) = w[n * sn + k * sk]
int gemm_tc_pack(const float* w, long long sn, long long sk, int N, int K, float* whi, float* wlo, cudaStream_t st) {
    int NT, Npad;
    gemm_tc_tiling(N, &NT, &Npad);
    gemm_tc_pack_kernel<<<64, 256, 0, st>>>(w, sn, sk, N, K, Npad, (int)align_up(K, GT_KC), whi, wlo);
    CGVP_LAUNCH_CHECK("gemm_tc_pack_kernel");
    return 0;
}

// C[M, N] = act(A W^T + bias) with W packed by gemm_tc_pack.  Returns 0 on success.
int gemm_tc(long long M, int N, int K, const float* A, long long lda, const float* whi, const float* wlo, const float* bias,
            float* C, long long ldc, int relu_cols, cudaStream_t st) {
    if (M <= 0 || N <= 0) return 0;
    GemmTcArgs a;
    memset(&a, 0, sizeof(a));
    a.M = M; a.N = N; a.K = K; a.Kpad = (int)align_up(K, GT_KC);
    gemm_tc_tiling(N, &a.NT, &a.Npad);
    a.A = A; a.lda = lda; a.whi = whi; a.wlo = wlo; a.bias = bias; a.C = C; a.ldc = ldc; a.relu_cols = relu_cols;
    const size_t smem = 128 + 2 * 8 * GT_LBO_A + 2 * 8 * (size_t)a.NT * 16 + 64;
    CGVP_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gemm_tc_kernel<<<dim3((unsigned)cdiv64(M, 128), (unsigned)(a.Npad / a.NT)), 128, smem, st>>>(a);
    CGVP_LAUNCH_CHECK("gemm_tc_kernel");
    return 0;
}

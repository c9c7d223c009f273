// Weight and bias gradient of a dense layer on the tensor cores, fp32-accurate ("3xTF32"), sm_100a:
//     dW[N, K] = dY[M, N]^T . X[M, K]        db[N] = sum_m dY[m, :]
// the autograd of the nn.Linear layers around the GVP encoder (residue / atom stacks, attention projections and
// feed-forward of models/joint_gnn.py:172-288,321-408), where M = all residues of the batch (~2 x 10^4) and N, K <= 256:
// a GEMM whose reduction runs over the ROWS of both operands.
//
// tcgen05 takes such operands as they lie in memory: both are "MN-major" (the non-reduced index is the contiguous
// one), staged in the one layout kind::tf32 accepts for that, SWIZZLE_128B_BASE32B -- rows of 32 floats, 4 reduction steps
// per 512-byte atom, 32-byte chunk c of step r at chunk c ^ (r & 3) -- so a warp copies a row-major tile with coalesced
// 16-byte loads and conflict-free 16-byte stores, no transposition.  Every fp32 value is split into hi = tf32(x) and lo = x - hi on the way;
// hi*hi + hi*lo + lo*hi accumulate in fp32 in TMEM (kind::tf32, M = 128 output rows, N = K_in columns, K = 8 rows per MMA).
// The rows are split over the CTAs (each adds its slab into its own [128, K] TMEM tile and writes a partial), the partials
// are summed in a fixed order by a second kernel: deterministic.  db rides along: the dY chunks pass through registers.
#include "cgvp_common.cuh"
#include "cgvp_tc.cuh"

#define WG_RB 16                          // rows (reduction steps) per stage = 2 MMA k-blocks (64 KB of stage buffers at K = 128:
                                          // small enough to share an SM with other streams' kernels)
#define WG_NT 512                         // threads per CTA: 16 warps share the split / store work of a stage
#define WG_NW (WG_NT / 32)
#define WG_GROUP_BYTES (WG_RB * 128)      // one 32-column group of a stage: [4-step atom][4 steps][128 B]

struct WgradArgs {
    long long M;
    int N, K;                             // N % 128 == 0, K % 32 == 0, K <= 256
    const float *dy, *x;
    float *pw, *pb;                       // partials [S][N][K], [S][N]
    long long rows_per_cta;
};

// chunk (r, c): 16 bytes = columns 4c .. 4c+3 of stage row r  ->  byte offset inside an operand tile
__device__ __forceinline__ int wg_off(int r, int c) {
    return (c >> 3) * WG_GROUP_BYTES + (r >> 2) * 512 + (r & 3) * 128 + (((((c & 7) >> 1) ^ (r & 3))) << 5) + ((c & 1) << 4);
}
__device__ __forceinline__ void wg_split_store(unsigned char* hi, unsigned char* lo, int off, float4 v) {
    float4 h, l;
    h.x = tcx::tf32_hi(v.x); h.y = tcx::tf32_hi(v.y); h.z = tcx::tf32_hi(v.z); h.w = tcx::tf32_hi(v.w);
    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
    *reinterpret_cast<float4*>(hi + off) = h;
    *reinterpret_cast<float4*>(lo + off) = l;
}

__global__ void __launch_bounds__(WG_NT) lin_wgrad_kernel(const __grid_constant__ WgradArgs a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = a.K, KG = K / 32, KC = K / 4;                  // column groups / 16-byte chunks per X row
    // two stage buffers, each [A_hi | A_lo | B_hi | B_lo]
    const int stage_bytes = 2 * 4 * WG_GROUP_BYTES + 2 * KG * WG_GROUP_BYTES;
    float* bsm = reinterpret_cast<float*>(smem + 2 * stage_bytes);             // [WG_NW warps][128] bias staging
    uint64_t* mbar = reinterpret_cast<uint64_t*>(bsm + WG_NW * 128);            // [2] one per buffer
    uint32_t* slot = reinterpret_cast<uint32_t*>(mbar + 2);
    const int tmem_cols = K <= 32 ? 32 : (K <= 64 ? 64 : (K <= 128 ? 128 : 256));
    if (tid == 0) { tcx::mbar_init(mbar, 1); tcx::mbar_init(mbar + 1, 1); tcx::fence_mbar_init(); }
    if (warp == 0) tcx::tmem_alloc(slot, tmem_cols);
    tcx::tc_fence_before();
    __syncthreads();
    tcx::tc_fence_after();
    const uint32_t tm0 = *slot;

    const int split = blockIdx.x, mt = blockIdx.y;               // row slab, 128-row tile of dW (= column tile of dY)
    const long long row_begin = (long long)split * a.rows_per_cta;
    const long long row_end = min(a.M, row_begin + a.rows_per_cta);
    const int nstages = (int)((row_end - row_begin + WG_RB - 1) / WG_RB);
    const float4* dy4 = reinterpret_cast<const float4*>(a.dy);
    const float4* x4 = reinterpret_cast<const float4*>(a.x);
    const long long dy_ld4 = a.N / 4;
    const int c_a = tid & 31;                                    // this thread's dY chunk column (fixed) ...
    const int r_a = tid >> 5;                                    // ... rows r_a, r_a + WG_NW, ... of a stage
    constexpr int NA = WG_RB / WG_NW;                            // dY chunks per thread per stage
    constexpr int NBMAX = (WG_RB * 64) / WG_NT;                  // X chunks per thread per stage at K = 256
    const int nchunk_b = WG_RB * KC, nb = (nchunk_b + WG_NT - 1) / WG_NT;
    float4 ra[NA], rb[NBMAX];
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    auto load_stage = [&](int st) {
        const long long r0 = row_begin + (long long)st * WG_RB;
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            const long long r = r0 + r_a + WG_NW * i;
            ra[i] = r < row_end ? __ldg(dy4 + r * dy_ld4 + mt * 32 + c_a) : zero4;
        }
#pragma unroll
        for (int i = 0; i < NBMAX; ++i) {
            const int q = tid + WG_NT * i;
            if (i < nb && q < nchunk_b) {
                const int r = q / KC, c = q - r * KC;
                const long long rr = r0 + r;
                rb[i] = rr < row_end ? __ldg(x4 + rr * KC + c) : zero4;
            }
        }
    };

    uint32_t phase[2] = {0, 0};
    if (nstages > 0) load_stage(0);
    for (int st = 0; st < nstages; ++st) {
        const int buf = st & 1;
        unsigned char* A_hi = smem + buf * stage_bytes;
        unsigned char* A_lo = A_hi + 4 * WG_GROUP_BYTES;
        unsigned char* B_hi = A_lo + 4 * WG_GROUP_BYTES;
        unsigned char* B_lo = B_hi + KG * WG_GROUP_BYTES;
        if (st >= 2) { tcx::mbar_wait(mbar + buf, phase[buf]); phase[buf] ^= 1; }   // the MMAs of stage st-2 have read this buffer
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            wg_split_store(A_hi, A_lo, wg_off(r_a + WG_NW * i, c_a), ra[i]);
            bsum.x += ra[i].x; bsum.y += ra[i].y; bsum.z += ra[i].z; bsum.w += ra[i].w;
        }
#pragma unroll
        for (int i = 0; i < NBMAX; ++i) {
            const int q = tid + WG_NT * i;
            if (i < nb && q < nchunk_b) {
                const int r = q / KC, c = q - r * KC;
                wg_split_store(B_hi, B_lo, wg_off(r, c), rb[i]);
            }
        }
        tcx::fence_proxy_async();
        tcx::tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tcx::tc_fence_after();
            const uint32_t id = tcx::idesc_tf32_mn(K);
            const uint32_t ah = tcx::smem_u32(A_hi), al = tcx::smem_u32(A_lo), bh = tcx::smem_u32(B_hi), bl = tcx::smem_u32(B_lo);
#pragma unroll
            for (int kb = 0; kb < WG_RB / 8; ++kb) {
                const uint64_t dah = tcx::smem_desc_mn32(ah + kb * 1024, WG_GROUP_BYTES, 512);
                const uint64_t dal = tcx::smem_desc_mn32(al + kb * 1024, WG_GROUP_BYTES, 512);
                const uint64_t dbh = tcx::smem_desc_mn32(bh + kb * 1024, WG_GROUP_BYTES, 512);
                const uint64_t dbl = tcx::smem_desc_mn32(bl + kb * 1024, WG_GROUP_BYTES, 512);
                tcx::mma_tf32(tm0, dal, dbh, id, (st > 0 || kb > 0) ? 1u : 0u);      // small terms first
                tcx::mma_tf32(tm0, dah, dbl, id, 1u);
                tcx::mma_tf32(tm0, dah, dbh, id, 1u);
            }
            tcx::mma_commit(mbar + buf);
        }
        if (st + 1 < nstages) load_stage(st + 1);                  // global loads of the next stage fly under the MMAs
    }
    // the last commit on each buffer covers every earlier MMA (commits complete in order)
    if (nstages > 0) { const int b = (nstages - 1) & 1; tcx::mbar_wait(mbar + b, phase[b]); phase[b] ^= 1; }
    if (nstages > 1) { const int b = (nstages - 2) & 1; tcx::mbar_wait(mbar + b, phase[b]); phase[b] ^= 1; }
    tcx::tc_fence_after();
    // partial dW tile: TMEM lane = dW row; warp w reads lanes 32 (w % 4) .. +31 and the column slices w / 4, w / 4 + 4, ...
    {
        const int lg = warp & 3;
        const uint32_t tm = tm0 + ((uint32_t)(lg * 32) << 16);
        float* out = a.pw + ((long long)split * a.N + mt * 128 + lg * 32 + lane) * K;
        for (int c0 = (warp >> 2) * 16; c0 < K; c0 += (WG_NW / 4) * 16) {
            float d[16];
            if (nstages > 0) {
                tcx::tmem_ld16(tm + c0, d);
                tcx::tmem_ld_wait(d);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) d[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) reinterpret_cast<float4*>(out + c0)[j] = make_float4(d[4 * j], d[4 * j + 1], d[4 * j + 2], d[4 * j + 3]);
        }
    }
    // partial db: the warps hold disjoint rows of the same 32 column chunks
    reinterpret_cast<float4*>(bsm + warp * 128)[lane] = bsum;
    tcx::tc_fence_before();
    __syncthreads();
    if (a.pb && tid < 128) {
        float sb = 0.f;
#pragma unroll
        for (int w2 = 0; w2 < WG_NW; ++w2) sb += bsm[w2 * 128 + tid];
        a.pb[(long long)split * a.N + mt * 128 + tid] = sb;
    }
    if (warp == 0) tcx::tmem_dealloc(tm0, tmem_cols);
}

// Sum of the S partials.  Block = 8 warps x 32 outputs: warp w sums partials w, w + 8, ... of its lane's output (coalesced,
// several loads in flight), the eight slices meet in shared memory in warp order: the association is fixed -> deterministic.
__global__ void __launch_bounds__(256) lin_wgrad_reduce_kernel(const float* __restrict__ pw, const float* __restrict__ pb, int S,
                                                                long long nk, int N, float* __restrict__ dw, float* __restrict__ db) {
    __shared__ float sm[8][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long i = (long long)blockIdx.x * 32 + lane;      // output index: [0, nk) = dW, [nk, nk + N) = db
    const bool is_w = i < nk, is_b = !is_w && db && i < nk + N;
    float s = 0.f;
    if (is_w) {
#pragma unroll 6
        for (int p = w; p < S; p += 8) s += __ldg(pw + (long long)p * nk + i);
    } else if (is_b) {
#pragma unroll 6
        for (int p = w; p < S; p += 8) s += __ldg(pb + (long long)p * N + (i - nk));
    }
    sm[w][lane] = s;
    __syncthreads();
    if (w == 0 && (is_w || is_b)) {
        float t = 0.f;
#pragma unroll
        for (int ww = 0; ww < 8; ++ww) t += sm[ww][lane];
        if (is_w) dw[i] = t; else db[i - nk] = t;
    }
}

static void wgrad_plan(int64_t M, int32_t N, int* S, long long* rows_per_cta) {
    const int mt = N / 128;
    long long s = cgvp_num_sms() / mt;
    if (s < 1) s = 1;
    const long long max_s = (M + 63) / 64;                         // at least 4 stages per CTA
    if (s > max_s) s = max_s;
    if (s < 1) s = 1;
    long long rpc = (M + s - 1) / s;
    rpc = (rpc + WG_RB - 1) / WG_RB * WG_RB;
    if (rpc < WG_RB) rpc = WG_RB;
    *rows_per_cta = rpc;
    *S = (int)((M + rpc - 1) / rpc);
    if (*S < 1) *S = 1;
}

extern "C" int32_t cgvp_linear_wgrad_supported(int64_t M, int32_t N, int32_t K) {
    return (M >= 1024 && N >= 128 && N % 128 == 0 && N <= 1024 && K >= 32 && K % 32 == 0 && K <= 256) ? 1 : 0;
}

extern "C" int64_t cgvp_linear_wgrad_workspace_bytes(int64_t M, int32_t N, int32_t K) {
    if (!cgvp_linear_wgrad_supported(M, N, K)) return -1;
    int S; long long rpc;
    wgrad_plan(M, N, &S, &rpc);
    return align_up((int64_t)S * N * K * 4, 256) + align_up((int64_t)S * N * 4, 256) + 256;
}

static bool wg_al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int32_t cgvp_linear_wgrad(const float* dy, const float* x, int64_t M, int32_t N, int32_t K, float* dw, float* db,
                                     void* ws, int64_t ws_bytes, void* stream) {
    CGVP_REQUIRE(cgvp_linear_wgrad_supported(M, N, K), "linear_wgrad: unsupported shape M=%lld N=%d K=%d", (long long)M, N, K);
    CGVP_REQUIRE(dy && x && dw && ws, "linear_wgrad: null argument");
    CGVP_REQUIRE(wg_al16(dy) && wg_al16(x) && wg_al16(dw), "linear_wgrad: buffers must be 16-byte aligned");
    CGVP_REQUIRE(ws_bytes >= cgvp_linear_wgrad_workspace_bytes(M, N, K), "linear_wgrad: workspace too small");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int S; long long rpc;
    wgrad_plan(M, N, &S, &rpc);
    char* b = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
    WgradArgs a;
    memset(&a, 0, sizeof(a));
    a.M = M; a.N = N; a.K = K; a.dy = dy; a.x = x; a.rows_per_cta = rpc;
    a.pw = reinterpret_cast<float*>(b);
    a.pb = reinterpret_cast<float*>(b + align_up((int64_t)S * N * K * 4, 256));
    const int KG = K / 32;
    const size_t smem = 1024 + 2 * (2 * 4 * WG_GROUP_BYTES + 2 * (size_t)KG * WG_GROUP_BYTES) + WG_NW * 128 * 4 + 64;
    CGVP_CUDA(cudaFuncSetAttribute(lin_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lin_wgrad_kernel<<<dim3((unsigned)S, (unsigned)(N / 128)), WG_NT, smem, st>>>(a);
    CGVP_LAUNCH_CHECK("lin_wgrad_kernel");
    const long long nk = (long long)N * K;
    lin_wgrad_reduce_kernel<<<(unsigned)cdiv64(nk + N, 32), 256, 0, st>>>(a.pw, a.pb, S, nk, N, dw, db);
    CGVP_LAUNCH_CHECK("lin_wgrad_reduce_kernel");
    return 0;
}

// ---- forward / input-gradient GEMM ------------------------------------------------------------------------------------------
//     Y[M, N] = A[M, K] . B[N, K]^T (+ bias),   B(n, k) = w[n * sn + k * sk]
// forward: A = x, B = weight (sn = K, sk = 1);  input gradient: A = dY, B = weight^T (sn = 1, sk = in_features).
// Persistent CTAs (one per SM): the CTA's NT-column slice of B is split (hi / lo tf32) ONCE into shared memory in the K-major
// core-matrix layout [k/4][n][4] and stays there; 128-row tiles of A stream through two 32-column staging buffers (global
// loads prefetched into registers while the previous chunk's MMAs run, split, stored), 12 kind::tf32 MMAs per chunk into one
// TMEM accumulator, epilogue (bias, store) per tile.  Same 3xTF32 arithmetic as gemm_tc.cu.
#define LG_KC 32
#define LG_LBO_A (2048 + 16)             // +16: the 8 k4 slices of a row land in 8 different bank groups

struct LinGemmArgs {
    long long M;
    int N, K, NT;                          // K % 32 == 0, N % NT == 0, NT % 16 == 0
    const float *A, *w, *bias;
    long long sn, sk;
    float* Y;
};

__global__ void __launch_bounds__(128, 1) lin_gemm_kernel(const __grid_constant__ LinGemmArgs a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int NT = a.NT, K = a.K, nchunks = K / LG_KC;
    const int b_lbo = NT * 16;                                            // bytes between k4 slices of B
    unsigned char* B_hi = smem;
    unsigned char* B_lo = B_hi + (size_t)(K / 4) * b_lbo;
    unsigned char* A_buf = B_lo + (size_t)(K / 4) * b_lbo;                // [2 buffers][hi, lo][8 k4 slices][LG_LBO_A]
    uint64_t* abar = reinterpret_cast<uint64_t*>(A_buf + 2 * 2 * 8 * LG_LBO_A);
    uint64_t* dbar = abar + 2;
    uint32_t* slot = reinterpret_cast<uint32_t*>(dbar + 1);
    const int tmem_cols = NT <= 32 ? 32 : (NT <= 64 ? 64 : (NT <= 128 ? 128 : 256));
    if (tid == 0) { tcx::mbar_init(abar, 1); tcx::mbar_init(abar + 1, 1); tcx::mbar_init(dbar, 1); tcx::fence_mbar_init(); }
    if (warp == 0) tcx::tmem_alloc(slot, tmem_cols);
    const int n0 = blockIdx.y * NT;
    // B slice -> shared memory, split.  Lanes run along the contiguous index of w (k for the forward use, n for the
    // transposed use); 8 independent loads in flight per thread.
    {
        const int total = NT * (K / 4), K4 = K / 4;
        const bool k_contig = a.sk == 1;
        for (int i0 = tid; i0 < total; i0 += 128 * 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u * 128;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < total) {
                    const int n = k_contig ? i / K4 : i % NT, k4 = k_contig ? i % K4 : i / NT;
                    const float* p = a.w + (long long)(n0 + n) * a.sn + (long long)(4 * k4) * a.sk;
                    if (k_contig) v[u] = __ldg(reinterpret_cast<const float4*>(p));
                    else { v[u].x = __ldg(p); v[u].y = __ldg(p + a.sk); v[u].z = __ldg(p + 2 * a.sk); v[u].w = __ldg(p + 3 * a.sk); }
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u * 128;
                if (i < total) {
                    const int n = k_contig ? i / K4 : i % NT, k4 = k_contig ? i % K4 : i / NT;
                    const float4 h = make_float4(tcx::tf32_hi(v[u].x), tcx::tf32_hi(v[u].y), tcx::tf32_hi(v[u].z), tcx::tf32_hi(v[u].w));
                    *reinterpret_cast<float4*>(B_hi + k4 * b_lbo + n * 16) = h;
                    *reinterpret_cast<float4*>(B_lo + k4 * b_lbo + n * 16) = make_float4(v[u].x - h.x, v[u].y - h.y, v[u].z - h.z, v[u].w - h.w);
                }
            }
        }
    }
    tcx::fence_proxy_async();
    tcx::tc_fence_before();
    __syncthreads();
    tcx::tc_fence_after();
    const uint32_t tm0 = *slot;
    const uint32_t tm = tm0 + ((uint32_t)(warp * 32) << 16);
    const uint32_t id = tcx::idesc_tf32(NT);
    const long long ntiles = (a.M + 127) / 128;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 regs[8];
    auto load_chunk = [&](long long tile, int c) {                        // 128 rows x 32 floats: lane -> (row, k4), k4 fastest
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int i = tid + q * 128, r = i >> 3, k4 = i & 7;
            const long long m = tile * 128 + r;
            regs[q] = m < a.M ? __ldg(reinterpret_cast<const float4*>(a.A + m * K + c * LG_KC) + k4) : zero4;
        }
    };
    uint32_t pa[2] = {0, 0}, pd = 0;
    int used[2] = {0, 0};                                                 // buffer holds operands of MMAs not yet known complete
    long long tile = blockIdx.x;
    if (tile < ntiles) load_chunk(tile, 0);
    for (; tile < ntiles; tile += gridDim.x) {
        for (int c = 0; c < nchunks; ++c) {
            const int buf = c & 1;
            unsigned char* Ah = A_buf + (size_t)buf * 2 * 8 * LG_LBO_A;
            unsigned char* Al = Ah + 8 * LG_LBO_A;
            if (used[buf]) { tcx::mbar_wait(abar + buf, pa[buf]); pa[buf] ^= 1; used[buf] = 0; }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int i = tid + q * 128, r = i >> 3, k4 = i & 7;
                const float4 x = regs[q];
                const float4 h = make_float4(tcx::tf32_hi(x.x), tcx::tf32_hi(x.y), tcx::tf32_hi(x.z), tcx::tf32_hi(x.w));
                *reinterpret_cast<float4*>(Ah + k4 * LG_LBO_A + r * 16) = h;
                *reinterpret_cast<float4*>(Al + k4 * LG_LBO_A + r * 16) = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
            }
            tcx::fence_proxy_async();
            tcx::tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tcx::tc_fence_after();
                const uint32_t ah = tcx::smem_u32(Ah), al = tcx::smem_u32(Al);
                const uint32_t bh = tcx::smem_u32(B_hi) + (uint32_t)(c * 8 * b_lbo), bl = tcx::smem_u32(B_lo) + (uint32_t)(c * 8 * b_lbo);
#pragma unroll
                for (int k8 = 0; k8 < LG_KC / 8; ++k8) {
                    const uint64_t dah = tcx::smem_desc(ah + k8 * 2 * LG_LBO_A, LG_LBO_A), dal = tcx::smem_desc(al + k8 * 2 * LG_LBO_A, LG_LBO_A);
                    const uint64_t dbh = tcx::smem_desc(bh + k8 * 2 * b_lbo, b_lbo), dbl = tcx::smem_desc(bl + k8 * 2 * b_lbo, b_lbo);
                    tcx::mma_tf32(tm0, dal, dbh, id, (c | k8) != 0);
                    tcx::mma_tf32(tm0, dah, dbl, id, 1);
                    tcx::mma_tf32(tm0, dah, dbh, id, 1);
                }
                tcx::mma_commit(abar + buf);
                if (c == nchunks - 1) tcx::mma_commit(dbar);
            }
            used[buf] = 1;
            // next chunk (possibly of the next tile) into registers while the tensor pipe works
            if (c + 1 < nchunks) load_chunk(tile, c + 1);
            else if (tile + gridDim.x < ntiles) load_chunk(tile + gridDim.x, 0);
        }
        tcx::mbar_wait(dbar, pd);
        pd ^= 1;
        tcx::tc_fence_after();
        const long long m = tile * 128 + tid;
        for (int cb = 0; cb < NT; cb += 16) {
            float d[16];
            tcx::tmem_ld16(tm + cb, d);
            tcx::tmem_ld_wait(d);
            if (m < a.M) {
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const int n = n0 + cb + 4 * j4;
                    float4 v = make_float4(d[4 * j4], d[4 * j4 + 1], d[4 * j4 + 2], d[4 * j4 + 3]);
                    if (a.bias) {
                        const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bias + n));
                        v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
                    }
                    *reinterpret_cast<float4*>(a.Y + m * a.N + n) = v;
                }
            }
        }
        tcx::tc_fence_before();
        __syncthreads();                                                  // TMEM reads done before the next tile overwrites it
        tcx::tc_fence_after();
    }
    // drain the per-buffer barriers (the last tile's commits) -- all covered by dbar, nothing to wait for
    tcx::tc_fence_before();
    __syncthreads();
    if (warp == 0) tcx::tmem_dealloc(tm0, tmem_cols);
}

static int lin_gemm_nt(int N, int K) {
    int nt = N < 128 ? N : 128;
    while (nt > 16 && ((long long)K * nt * 8 > 131072 || N % nt != 0)) nt -= 16;
    return nt;
}

extern "C" int32_t cgvp_linear_gemm_supported(int64_t M, int32_t N, int32_t K) {
    if (!(M >= 1024 && N >= 16 && N % 16 == 0 && N <= 4096 && K >= 32 && K % 32 == 0 && K <= 1024)) return 0;
    const int nt = lin_gemm_nt(N, K);
    return (nt >= 16 && N % nt == 0 && (long long)K * nt * 8 <= 131072) ? 1 : 0;
}

static int lin_gemm_launch(const float* A, const float* w, long long sn, long long sk, const float* bias, int64_t M, int32_t N,
                           int32_t K, float* Y, cudaStream_t st) {
    CGVP_REQUIRE(cgvp_linear_gemm_supported(M, N, K), "linear_gemm: unsupported shape M=%lld N=%d K=%d", (long long)M, N, K);
    CGVP_REQUIRE(A && w && Y, "linear_gemm: null argument");
    CGVP_REQUIRE(wg_al16(A) && wg_al16(Y) && wg_al16(bias) && wg_al16(w), "linear_gemm: buffers must be 16-byte aligned");
    LinGemmArgs a;
    memset(&a, 0, sizeof(a));
    a.M = M; a.N = N; a.K = K; a.NT = lin_gemm_nt(N, K); a.A = A; a.w = w; a.bias = bias; a.sn = sn; a.sk = sk; a.Y = Y;
    const size_t smem = 128 + 2 * (size_t)(K / 4) * a.NT * 16 + 2 * 2 * 8 * LG_LBO_A + 64;
    CGVP_CUDA(cudaFuncSetAttribute(lin_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ntile_n = N / a.NT;
    long long gx = cgvp_num_sms() / ntile_n;
    if (gx < 1) gx = 1;
    const long long mtiles = (M + 127) / 128;
    if (gx > mtiles) gx = mtiles;
    lin_gemm_kernel<<<dim3((unsigned)gx, (unsigned)ntile_n), 128, smem, st>>>(a);
    CGVP_LAUNCH_CHECK("lin_gemm_kernel");
    return 0;
}

// y[M, N] = x[M, K] w[N, K]^T + bias         (nn.Linear forward)
extern "C" int32_t cgvp_linear_fwd(const float* x, const float* w, const float* bias, int64_t M, int32_t N, int32_t K, float* y,
                                   void* stream) {
    return lin_gemm_launch(x, w, K, 1, bias, M, N, K, y, reinterpret_cast<cudaStream_t>(stream));
}
// dx[M, K] = dy[M, N] w[N, K]                (nn.Linear input gradient)
extern "C" int32_t cgvp_linear_dgrad(const float* dy, const float* w, int64_t M, int32_t N, int32_t K, float* dx, void* stream) {
    return lin_gemm_launch(dy, w, 1, K, nullptr, M, K, N, dx, reinterpret_cast<cudaStream_t>(stream));
}

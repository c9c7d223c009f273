// Tensor-core fused GVPConv forward (tcgen05 + TMEM, sm_100a) for wide node dims -- BASELINE config 5:
// nodes (100,16), edges (32,1), message chain GVP(relu,gate) -> GVP(relu,gate) -> GVP(none,gate).
// Replaces GVPConv.forward/message + PyG propagate (models/gvp_layers.py:291-308) like conv.cu / conv_reg.cu.
//
// Per CTA (one per SM, persistent): two warpgroups, each owning one 128-edge tile at a time (thread = edge = TMEM
// lane), ping-pong on the tensor pipe.  Every W_h / W_s / W_mu(wv) / gate projection of the three message GVPs is a
// tcgen05.mma (kind::f16, bf16 operands, fp32 accumulation in TMEM, M = 128 edges):
//     A operand  = the tile's activations, written by the owning threads as bf16 into the canonical K-major
//                  no-swizzle layout  [k/8][row][8]  (8x16-byte core matrices, SBO = 128 B, LBO = 2048 B);
//     B operand  = weights, pre-packed once per call into the same layout and staged into shared memory with
//                  cp.async.bulk (TMA engine) on an mbarrier; they stay resident for the whole kernel;
//     D          = TMEM columns, read back with tcgen05.ld.32x32b (one row per thread) for the fused epilogues:
//                  vector norms (:153), bias rows, ReLU (:172), sigmoid gate (:158-163).
// Linearity before the gather: the s_j / s_i blocks of W_s and the V_j / V_i blocks of W_h of the FIRST message GVP
// act on per-node rows, so they are applied once per node (tc_node_proj_kernel) and the per-edge GEMM K drops from
// 265 to 66; the projected rows are gathered and added in the epilogue.
// Aggregation: deterministic segmented sum over the sorted targets (per-tile pieces + conv_fixup_kernel).
// Accuracy: bf16 operands -> scale-relative error ~2e-3 (north star allows <= 1e-2 with tensor cores); the fp32
// paths (conv_reg.cu / conv.cu) remain the default -- see cgvp_set_tensor_cores().
#include <cuda_bf16.h>

#include "cgvp_reg.cuh"

using namespace cgvpr;

constexpr CGVP_HD inline int pad16(int x) { return (x + 15) / 16 * 16; }

template <int NS_, int NV_, int ES_, int EV_>
struct TcSpec {
    static constexpr int NS = NS_, NV = NV_, ES = ES_, EV = EV_;
    static constexpr int H0 = 2 * NV + EV;                       // hidden vector channels of message GVP 0
    static constexpr int HQ = pad4(H0);                          // row pitch of the projected vector tables
    static constexpr int SI0 = 2 * NS + ES, KSD0 = SI0 + H0;     // ws input width of GVP 0 (before the ones column)
    static constexpr int SOP = pad4(NS), VOP = pad4(NV), HP0 = pad4(H0), HP1 = pad4(NV);
    // padded GEMM shapes (bf16: K multiple of 16; M = 128 needs N multiple of 16)
    static constexpr int N_S = pad16(NS), N_V = pad16(NV);
    static constexpr int K_S0 = pad16(ES + H0 + 1), K_V0 = pad16(H0), K_G = pad16(NS + 1), K_H = pad16(NV), K_S1 = pad16(NS + NV + 1);
    // weight arena (bytes, bf16, [k/8][n][8] blocks)
    static constexpr int W_S0 = 0, W_V0 = W_S0 + N_S * K_S0 * 2, W_G0 = W_V0 + N_V * K_V0 * 2, W_ST1 = W_G0 + N_V * K_G * 2;
    static constexpr int W_H = 0, W_S = W_H + N_V * K_H * 2, W_V = W_S + N_S * K_S1 * 2, W_G = W_V + N_V * K_H * 2, W_STAGE = W_G + N_V * K_G * 2;
    static constexpr int W_BYTES = W_ST1 + 2 * W_STAGE;
    static constexpr int WHE_FLOATS = pad4(max1(EV) * HQ);       // fp32: wh columns of the edge vector channels
    // TMEM columns per warpgroup
    static constexpr int C_VH = 0, C_S = C_VH + 3 * N_V, C_VO = C_S + N_S, C_G = C_VO + 3 * N_V, C_END = C_G + N_V;
    static_assert(C_END <= 256, "TMEM columns per warpgroup");
    // activation tile region per warpgroup (bytes)
    static constexpr int A_S0 = 0, A_VH0 = A_S0 + (K_S0 / 8) * 2048;
    static constexpr int A_S1 = 0, A_V1 = A_S1 + (K_S1 / 8) * 2048;
    static constexpr int CH = NS + 3 * NV, CHH = (CH + 1) / 2;   // message channels; reduced in two halves
    static constexpr int TILE_BYTES = imax(imax(A_VH0 + 3 * (K_V0 / 8) * 2048, A_V1 + 3 * (K_H / 8) * 2048), (int)align_up(CHH * 129 * 4, 16));
    static constexpr int WG_BYTES = TILE_BYTES + 4 * 128 * 4 + 64;   // + src/dst/eid/scale + mbarrier
    static constexpr size_t smem_bytes() { return 1024 + (size_t)W_BYTES + WHE_FLOATS * 4 + 2 * (size_t)WG_BYTES + 64; }
    static bool matches(const CgvpConvDesc& d) {
        using G0 = GvpC<SI0, H0, NS, NV, H0, CGVP_ACT_RELU, CGVP_ACT_NONE, 1>;
        using G1 = GvpC<NS, NV, NS, NV, NV, CGVP_ACT_RELU, CGVP_ACT_NONE, 1>;
        using G2 = GvpC<NS, NV, NS, NV, NV, CGVP_ACT_NONE, CGVP_ACT_NONE, 1>;
        return d.ns == NS && d.nv == NV && d.es == ES && d.ev == EV && d.n_gvp == 3 && G0::matches(d.gvp[0]) &&
               G1::matches(d.gvp[1]) && G2::matches(d.gvp[2]);
    }
};

struct TcArgs {
    long long E, N;
    int ntiles, mean, edge_sorted;
    const int *perm, *src, *dst, *rowptr;
    const float *e_s, *e_v;
    const float *psj, *psi, *pvj, *pvi;     // per-node projections [N][NS], [N][NS], [N][3][HQ], [N][3][HQ]
    const unsigned char* wtc;               // bf16 weight arena
    const float* whe;                       // [EV][HQ]
    float *out_s, *out_v, *part_head, *part_tail;
};

// ---- PTX wrappers ---------------------------------------------------------------------------------------------------
namespace tcx {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done, spins = 0;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 26)) __trap();      // a lost arrival must fail loudly, never hang the GPU
    } while (!done);
}
// 1-D bulk copy global -> shared through the TMA engine, completion counted on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wg_sync(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, int cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, int cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// K-major, no swizzle: core matrix = 8 rows x 16 bytes, rows contiguous (SBO = 128 B), k-chunks LBO bytes apart
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128
__device__ __forceinline__ constexpr uint32_t idesc_bf16(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 16 consecutive fp32 columns of this thread's TMEM lane.  The loads are asynchronous: issue a batch, then ONE
// tmem_ld_wait() before the registers are read (the "+f" constraints of the wait keep the compiler from moving uses up).
__device__ __forceinline__ void tmem_ld16(uint32_t addr, float* d) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]), "=f"(d[4]), "=f"(d[5]), "=f"(d[6]), "=f"(d[7]), "=f"(d[8]),
                   "=f"(d[9]), "=f"(d[10]), "=f"(d[11]), "=f"(d[12]), "=f"(d[13]), "=f"(d[14]), "=f"(d[15])
                 : "r"(addr));
}
template <int N>
__device__ __forceinline__ void tmem_ld_wait(float (&d)[N]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < N; ++i) asm volatile("" : "+f"(d[i]));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
// bf16-operand mode: approximate SFU maths is far inside the mode's error budget
__device__ __forceinline__ float fast_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
// 8 consecutive K values of one row -> one 16-byte core-matrix row
__device__ __forceinline__ void put8(unsigned char* region, int chunk, int row, const float (&v)[8]) {
    uint4 q;
    q.x = pack_bf16(v[0], v[1]); q.y = pack_bf16(v[2], v[3]); q.z = pack_bf16(v[4], v[5]); q.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(region + chunk * 2048 + row * 16) = q;
}
}  // namespace tcx
using namespace tcx;

// D[128 x N] (+)= A[128 x K] . B[N x K]^T, K in steps of 16 (two 16-byte k-chunks per instruction)
template <int N, int K>
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr) {
    constexpr uint32_t id = idesc_bf16(N);
#pragma unroll
    for (int k = 0; k < K / 16; ++k)
        mma_bf16(tmem_d, smem_desc(a_addr + k * 2 * 2048, 2048), smem_desc(b_addr + k * 2 * (N * 16), N * 16), id, k > 0);
}

// ---- weight pre-packing -------------------------------------------------------------------------------------------------
// Generic fp32 packed blocks (cgvp_common.cuh) -> bf16 [k/8][n][8] blocks of the padded GEMM shapes.
template <class S>
__global__ void tc_pack_kernel(const float* __restrict__ w0, const float* __restrict__ w1, const float* __restrict__ w2,
                               __nv_bfloat16* __restrict__ out, float* __restrict__ whe) {
    using G0 = GvpC<S::SI0, S::H0, S::NS, S::NV, S::H0, 1, 0, 1>;
    using G1 = GvpC<S::NS, S::NV, S::NS, S::NV, S::NV, 1, 0, 1>;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < S::WHE_FLOATS) {
        const int c = i / S::HQ, o = i % S::HQ;
        whe[i] = (c < S::EV && o < S::H0) ? w0[G0::O_WH_T + (S::NV + c) * S::HP0 + o] : 0.f;
    }
    if (i >= S::W_BYTES / 2) return;
    const int byte = 2 * i;
    float v = 0.f;
    auto blk = [&](int off, int npad, int& n, int& k) {    // element index inside a [k/8][npad][8] block
        const int j = (byte - off) / 2;
        k = (j / (npad * 8)) * 8 + (j & 7);
        n = (j >> 3) % npad;
    };
    int n, k;
    if (byte < S::W_V0) {                                  // ws of GVP 0, edge part: [e_s ; vn ; 1]
        blk(S::W_S0, S::N_S, n, k);
        if (n < S::NS) {
            if (k < S::ES) v = w0[G0::O_WS_T + (S::NS + k) * S::SOP + n];
            else if (k < S::ES + S::H0) v = w0[G0::O_WS_T + (S::SI0 + (k - S::ES)) * S::SOP + n];
            else if (k == S::ES + S::H0) v = w0[G0::O_WS_T + S::KSD0 * S::SOP + n];
        }
    } else if (byte < S::W_G0) {                           // wv of GVP 0
        blk(S::W_V0, S::N_V, n, k);
        if (n < S::NV && k < S::H0) v = w0[G0::O_WV_T + k * S::VOP + n];
    } else if (byte < S::W_ST1) {                          // gate of GVP 0: [s' ; 1]
        blk(S::W_G0, S::N_V, n, k);
        if (n < S::NV && k <= S::NS) v = w0[G0::O_WSV_T + k * S::VOP + n];
    } else {
        const int st = (byte - S::W_ST1) / S::W_STAGE;
        const int base = S::W_ST1 + st * S::W_STAGE;
        const float* w = st == 0 ? w1 : w2;
        const int b = byte - base;
        if (b < S::W_S) {                                  // wh
            blk(base + S::W_H, S::N_V, n, k);
            if (n < S::NV && k < S::NV) v = w[G1::O_WH_T + k * S::HP1 + n];
        } else if (b < S::W_V) {                           // ws: [s ; vn ; 1]
            blk(base + S::W_S, S::N_S, n, k);
            if (n < S::NS && k <= S::NS + S::NV) v = w[G1::O_WS_T + k * S::SOP + n];
        } else if (b < S::W_G) {                           // wv
            blk(base + S::W_V, S::N_V, n, k);
            if (n < S::NV && k < S::NV) v = w[G1::O_WV_T + k * S::VOP + n];
        } else {                                           // gate
            blk(base + S::W_G, S::N_V, n, k);
            if (n < S::NV && k <= S::NS) v = w[G1::O_WSV_T + k * S::VOP + n];
        }
    }
    out[i] = __float2bfloat16_rn(v);
}

// ---- per-node projections of message GVP 0 (fp32) -------------------------------------------------------------------------
//   psj[n] = W_s[:, s_j block] x_s[n]      psi[n] = W_s[:, s_i block] x_s[n]
//   pvj[n][p] = W_h[:, V_j block] x_V[n][:, p]   pvi likewise with the V_i block
template <class S>
__global__ void __launch_bounds__(256) tc_node_proj_kernel(long long N, const float* __restrict__ x_s, const float* __restrict__ x_v,
                                                            const float* __restrict__ w0, float* __restrict__ psj,
                                                            float* __restrict__ psi, float* __restrict__ pvj, float* __restrict__ pvi) {
    using G0 = GvpC<S::SI0, S::H0, S::NS, S::NV, S::H0, 1, 0, 1>;
    constexpr int NB = 8;                                  // nodes per pass
    __shared__ float xs[S::NS][NB];
    __shared__ float xv[3 * S::NV][NB];                    // [p * NV + c][node]
    const int t = threadIdx.x;
    for (long long n0 = (long long)blockIdx.x * NB; n0 < N; n0 += (long long)gridDim.x * NB) {
        __syncthreads();
        for (int i = t; i < S::NS * NB; i += blockDim.x) {
            const int nn = i / S::NS, k = i % S::NS;
            xs[k][nn] = n0 + nn < N ? x_s[(n0 + nn) * S::NS + k] : 0.f;
        }
        for (int i = t; i < 3 * S::NV * NB; i += blockDim.x) {
            const int nn = i / (3 * S::NV), j = i % (3 * S::NV), c = j / 3, p = j % 3;
            xv[p * S::NV + c][nn] = n0 + nn < N ? x_v[(n0 + nn) * 3 * S::NV + j] : 0.f;
        }
        __syncthreads();
        float acc[NB];
        if (t < 2 * S::NS) {                               // scalar projections: side = t / NS, output o = t % NS
            const int side = t / S::NS, o = t % S::NS;
            const float* w = w0 + G0::O_WS_T + (side ? (S::NS + S::ES) : 0) * S::SOP + o;
#pragma unroll
            for (int j = 0; j < NB; ++j) acc[j] = 0.f;
            for (int k = 0; k < S::NS; ++k) {
                const float wk = __ldg(w + k * S::SOP);
#pragma unroll
                for (int j = 0; j < NB; ++j) acc[j] = fmaf(xs[k][j], wk, acc[j]);
            }
            float* out = side ? psi : psj;
#pragma unroll
            for (int j = 0; j < NB; ++j)
                if (n0 + j < N) out[(n0 + j) * S::NS + o] = acc[j];
        }
        for (int q = t; q < 2 * 3 * S::HQ; q += blockDim.x) {   // vector projections: (side, plane, o)
            const int side = q / (3 * S::HQ), r = q % (3 * S::HQ), p = r / S::HQ, o = r % S::HQ;
#pragma unroll
            for (int j = 0; j < NB; ++j) acc[j] = 0.f;
            if (o < S::H0) {
                const float* w = w0 + G0::O_WH_T + (side ? (S::NV + S::EV) : 0) * S::HP0 + o;
                for (int c = 0; c < S::NV; ++c) {
                    const float wc = __ldg(w + c * S::HP0);
#pragma unroll
                    for (int j = 0; j < NB; ++j) acc[j] = fmaf(xv[p * S::NV + c][j], wc, acc[j]);
                }
            }
            float* out = side ? pvi : pvj;
#pragma unroll
            for (int j = 0; j < NB; ++j)
                if (n0 + j < N) out[((n0 + j) * 3 + p) * S::HQ + o] = acc[j];
        }
    }
}

// ---- main kernel ----------------------------------------------------------------------------------------------------------
template <int W>
__device__ __forceinline__ void ld_row(const float* __restrict__ p, float* d) {     // W % 4 == 0, 16-byte aligned
#pragma unroll
    for (int i = 0; i < W / 4; ++i) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
        d[4 * i] = t.x; d[4 * i + 1] = t.y; d[4 * i + 2] = t.z; d[4 * i + 3] = t.w;
    }
}
template <int W>
__device__ __forceinline__ void add_row(const float* __restrict__ p, float* d) {
#pragma unroll
    for (int i = 0; i < W / 4; ++i) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
        d[4 * i] += t.x; d[4 * i + 1] += t.y; d[4 * i + 2] += t.z; d[4 * i + 3] += t.w;
    }
}

// epilogue shared by the three stages: s' (pre-activation, already complete) and Vo / gate in TMEM -> (s_out, V_out)
template <class S, bool RELU>
__device__ __forceinline__ void gate_and_finish(unsigned char* tile, uint32_t wsm_gate, uint32_t tm, uint32_t tm_wg, int wg, int row, bool leader,
                                                uint64_t* bar, uint32_t& phase, float (&s)[S::N_S], float (&v)[3][S::N_V]) {
    // gate input = s' (vector_act is None), + ones column for the bias
#pragma unroll
    for (int c = 0; c < S::K_G / 8; ++c) {
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = 8 * c + j;
            a[j] = k < S::NS ? s[k < S::NS ? k : 0] : (k == S::NS ? 1.f : 0.f);
        }
        put8(tile + S::A_S1, c, row, a);
    }
    fence_proxy_async();
    tc_fence_before();
    wg_sync(wg);
    if (leader) {
        tc_fence_after();
        issue_gemm<S::N_V, S::K_G>(tm_wg + S::C_G, smem_u32(tile + S::A_S1), wsm_gate);
        mma_commit(bar);
    }
    if (RELU) {
#pragma unroll
        for (int k = 0; k < S::NS; ++k) s[k] = fmaxf(s[k], 0.f);                        // :172-173 (after the gate input was taken)
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    float g[S::N_V];
#pragma unroll
    for (int c = 0; c < S::N_V / 16; ++c) tmem_ld16(tm + S::C_G + 16 * c, g + 16 * c);
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int c = 0; c < S::N_V / 16; ++c) tmem_ld16(tm + S::C_VO + p * S::N_V + 16 * c, v[p] + 16 * c);
    tmem_ld_wait(g); tmem_ld_wait(v[0]); tmem_ld_wait(v[1]); tmem_ld_wait(v[2]);
#pragma unroll
    for (int c = 0; c < S::NV; ++c) g[c] = fast_sigmoid(g[c]);
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int c = 0; c < S::NV; ++c) v[p][c] *= g[c];                                 // :163
}

template <class S>
__global__ void __launch_bounds__(256, 1) conv_tc_fwd_kernel(const __grid_constant__ TcArgs a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* wsm = smem;
    float* whe = reinterpret_cast<float*>(smem + S::W_BYTES);
    unsigned char* wgbase = smem + S::W_BYTES + S::WHE_FLOATS * 4;
    const int tid = threadIdx.x, wg = tid >> 7, row = tid & 127, warp = tid >> 5;
    unsigned char* tile = wgbase + wg * S::WG_BYTES;
    int* isrc = reinterpret_cast<int*>(tile + S::TILE_BYTES);
    int* idst = isrc + 128;
    int* ieid = idst + 128;
    uint64_t* bar = reinterpret_cast<uint64_t*>(ieid + 2 * 128);
    uint64_t* wbar = reinterpret_cast<uint64_t*>(wgbase + 2 * S::WG_BYTES);
    uint32_t* slot = reinterpret_cast<uint32_t*>(wbar + 1);

    if (tid == 0) {
        mbar_init(wbar, 1);
        mbar_init(reinterpret_cast<uint64_t*>(wgbase + S::TILE_BYTES + 4 * 128 * 4), 1);
        mbar_init(reinterpret_cast<uint64_t*>(wgbase + S::WG_BYTES + S::TILE_BYTES + 4 * 128 * 4), 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(slot, 512);
    for (int i = tid; i < S::WHE_FLOATS; i += blockDim.x) whe[i] = a.whe[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {                                        // weights: TMA bulk copies, resident for the whole kernel
        mbar_expect_tx(wbar, S::W_BYTES);
        constexpr int CHUNK = 16384;
        for (int off = 0; off < S::W_BYTES; off += CHUNK)
            bulk_g2s(wsm + off, a.wtc + off, (uint32_t)(S::W_BYTES - off < CHUNK ? S::W_BYTES - off : CHUNK), wbar);
    }
    const uint32_t tm = *slot + (uint32_t)(wg * 256) + ((uint32_t)((warp & 3) * 32) << 16);   // this thread's lane, this WG's columns
    const uint32_t tm_wg = *slot + (uint32_t)(wg * 256);                                       // MMA destination (lane 0)
    mbar_wait(wbar, 0);
    uint32_t phase = 0;
    const bool leader = row == 0;
    const uint32_t w0s = smem_u32(wsm);

    for (int t = blockIdx.x * 2 + wg; t < a.ntiles; t += gridDim.x * 2) {
        const long long p0 = (long long)t * 128;
        const int rv = (int)min(128LL, a.E - p0);
        const long long p = row < rv ? p0 + row : p0;       // idle rows replay the first edge (never stored)
        const int src = __ldg(a.src + p), dst = __ldg(a.dst + p);
        const long long eid = a.edge_sorted ? p : (long long)__ldg(a.perm + p);
        wg_sync(wg);                                        // previous tile's reduce is done with the tile region / index arrays
        isrc[row] = src; idst[row] = dst; ieid[row] = (int)eid;

        float s[S::N_S], v[3][S::N_V];
#pragma unroll
        for (int k = S::NS; k < S::N_S; ++k) s[k] = 0.f;
        // ================= message GVP 0 =================
        {
            // Vh = W_h [V_j ; e_V ; V_i] = pvj[src] + pvi[dst] + sum_c e_V[c] (x) whe[c]                       :152
            float vh[3][S::HQ];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                ld_row<S::HQ>(a.pvj + ((long long)src * 3 + q) * S::HQ, vh[q]);
                add_row<S::HQ>(a.pvi + ((long long)dst * 3 + q) * S::HQ, vh[q]);
            }
            if constexpr (S::EV > 0) {
#pragma unroll
                for (int c = 0; c < S::EV; ++c) {
                    const float ex = __ldg(a.e_v + (eid * S::EV + c) * 3), ey = __ldg(a.e_v + (eid * S::EV + c) * 3 + 1),
                                ez = __ldg(a.e_v + (eid * S::EV + c) * 3 + 2);
#pragma unroll
                    for (int o = 0; o < S::H0; ++o) {
                        const float w = whe[c * S::HQ + o];
                        vh[0][o] = fmaf(ex, w, vh[0][o]); vh[1][o] = fmaf(ey, w, vh[1][o]); vh[2][o] = fmaf(ez, w, vh[2][o]);
                    }
                }
            }
            // A operands: [e_s ; vn ; 1] for W_s (edge part) and Vh (3 planes) for W_mu
            float es[S::ES];
            ld_row<S::ES>(a.e_s + eid * S::ES, es);
#pragma unroll
            for (int c = 0; c < S::K_S0 / 8; ++c) {
                float q8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int k = 8 * c + j;
                    if (k < S::ES) q8[j] = es[k < S::ES ? k : 0];
                    else if (k < S::ES + S::H0) {
                        const int o = k - S::ES < S::H0 ? (k - S::ES >= 0 ? k - S::ES : 0) : 0;
                        q8[j] = fast_sqrt(fmaxf(vh[0][o] * vh[0][o] + vh[1][o] * vh[1][o] + vh[2][o] * vh[2][o], CGVP_EPS));   // :153
                    } else q8[j] = k == S::ES + S::H0 ? 1.f : 0.f;
                }
                put8(tile + S::A_S0, c, row, q8);
            }
#pragma unroll
            for (int q = 0; q < 3; ++q)
#pragma unroll
                for (int c = 0; c < S::K_V0 / 8; ++c) {
                    float q8[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) { const int o = 8 * c + j; q8[j] = o < S::H0 ? vh[q][o < S::H0 ? o : 0] : 0.f; }
                    put8(tile + S::A_VH0 + q * (S::K_V0 / 8) * 2048, c, row, q8);
                }
            fence_proxy_async();
            tc_fence_before();
            wg_sync(wg);
            if (leader) {
                tc_fence_after();
                issue_gemm<S::N_S, S::K_S0>(tm_wg + S::C_S, smem_u32(tile + S::A_S0), w0s + S::W_S0);
#pragma unroll
                for (int q = 0; q < 3; ++q)
                    issue_gemm<S::N_V, S::K_V0>(tm_wg + S::C_VO + q * S::N_V, smem_u32(tile + S::A_VH0 + q * (S::K_V0 / 8) * 2048), w0s + S::W_V0);
                mma_commit(bar);
            }
            // while the tensor pipe works: the node-projected part of s'
            ld_row<S::NS>(a.psj + (long long)src * S::NS, s);
            add_row<S::NS>(a.psi + (long long)dst * S::NS, s);
            mbar_wait(bar, phase);
            phase ^= 1;
            tc_fence_after();
            {
                constexpr int NB = S::N_S / 16, B0 = (NB + 1) / 2;      // two batches keep the temporaries small
                float d0[16 * B0], d1[16 * (NB - B0 > 0 ? NB - B0 : 1)];
#pragma unroll
                for (int c = 0; c < B0; ++c) tmem_ld16(tm + S::C_S + 16 * c, d0 + 16 * c);
#pragma unroll
                for (int c = B0; c < NB; ++c) tmem_ld16(tm + S::C_S + 16 * c, d1 + 16 * (c - B0));
                tmem_ld_wait(d0); tmem_ld_wait(d1);
#pragma unroll
                for (int j = 0; j < 16 * B0; ++j)
                    if (j < S::NS) s[j < S::NS ? j : 0] += d0[j];
#pragma unroll
                for (int j = 16 * B0; j < 16 * NB; ++j)
                    if (j < S::NS) s[j < S::NS ? j : 0] += d1[j - 16 * B0];
            }
            gate_and_finish<S, true>(tile, w0s + S::W_G0, tm, tm_wg, wg, row, leader, bar, phase, s, v);
        }
        // ================= message GVPs 1 and 2 =================
#pragma unroll
        for (int st = 0; st < 2; ++st) {
            const uint32_t wst = w0s + S::W_ST1 + st * S::W_STAGE;
            // Vh = W_h V                                                                                          :152
#pragma unroll
            for (int q = 0; q < 3; ++q)
#pragma unroll
                for (int c = 0; c < S::K_H / 8; ++c) {
                    float q8[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) { const int o = 8 * c + j; q8[j] = o < S::NV ? v[q][o < S::NV ? o : 0] : 0.f; }
                    put8(tile + S::A_V1 + q * (S::K_H / 8) * 2048, c, row, q8);
                }
            fence_proxy_async();
            tc_fence_before();
            wg_sync(wg);
            if (leader) {
                tc_fence_after();
#pragma unroll
                for (int q = 0; q < 3; ++q)
                    issue_gemm<S::N_V, S::K_H>(tm_wg + S::C_VH + q * S::N_V, smem_u32(tile + S::A_V1 + q * (S::K_H / 8) * 2048), wst + S::W_H);
                mma_commit(bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1;
            tc_fence_after();
            float vn[S::N_V];
            {
                float vh[3][S::N_V];
#pragma unroll
                for (int q = 0; q < 3; ++q)
#pragma unroll
                    for (int c = 0; c < S::N_V / 16; ++c) tmem_ld16(tm + S::C_VH + q * S::N_V + 16 * c, vh[q] + 16 * c);
                tmem_ld_wait(vh[0]); tmem_ld_wait(vh[1]); tmem_ld_wait(vh[2]);
#pragma unroll
                for (int o = 0; o < S::N_V; ++o)
                    vn[o] = fast_sqrt(fmaxf(vh[0][o] * vh[0][o] + vh[1][o] * vh[1][o] + vh[2][o] * vh[2][o], CGVP_EPS));     // :153
                // Vh back as the A operand of W_mu (the MMAs that read V from this region have completed)
#pragma unroll
                for (int q = 0; q < 3; ++q)
#pragma unroll
                    for (int c = 0; c < S::K_H / 8; ++c) {
                        float q8[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) { const int o = 8 * c + j; q8[j] = o < S::NV ? vh[q][o < S::NV ? o : 0] : 0.f; }
                        put8(tile + S::A_V1 + q * (S::K_H / 8) * 2048, c, row, q8);
                    }
            }
            // [s ; vn ; 1]
#pragma unroll
            for (int c = 0; c < S::K_S1 / 8; ++c) {
                float q8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int k = 8 * c + j;
                    if (k < S::NS) q8[j] = s[k < S::NS ? k : 0];
                    else if (k < S::NS + S::NV) q8[j] = vn[k - S::NS < S::NV ? (k - S::NS >= 0 ? k - S::NS : 0) : 0];
                    else q8[j] = k == S::NS + S::NV ? 1.f : 0.f;
                }
                put8(tile + S::A_S1, c, row, q8);
            }
            fence_proxy_async();
            tc_fence_before();
            wg_sync(wg);
            if (leader) {
                tc_fence_after();
                issue_gemm<S::N_S, S::K_S1>(tm_wg + S::C_S, smem_u32(tile + S::A_S1), wst + S::W_S);
#pragma unroll
                for (int q = 0; q < 3; ++q)
                    issue_gemm<S::N_V, S::K_H>(tm_wg + S::C_VO + q * S::N_V, smem_u32(tile + S::A_V1 + q * (S::K_H / 8) * 2048), wst + S::W_V);
                mma_commit(bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1;
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < S::N_S / 16; ++c) tmem_ld16(tm + S::C_S + 16 * c, s + 16 * c);
            tmem_ld_wait(s);
            if (st == 0) gate_and_finish<S, true>(tile, wst + S::W_G, tm, tm_wg, wg, row, leader, bar, phase, s, v);
            else gate_and_finish<S, false>(tile, wst + S::W_G, tm, tm_wg, wg, row, leader, bar, phase, s, v);
        }
        // ================= aggregation: segmented sum over the sorted targets, two channel halves =================
        tc_fence_before();
        float* M = reinterpret_cast<float*>(tile);
        const int n_first = idst[0], n_last = idst[rv - 1];
        const long long p1 = p0 + rv;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            wg_sync(wg);                                    // tile region free (MMAs done; previous half consumed)
            const int ch0 = half * S::CHH;
#pragma unroll
            for (int c = 0; c < S::CHH; ++c) {
                const int ch = ch0 + c;
                if (ch < S::CH) {
                    float val;
                    if (ch < S::NS) val = s[ch < S::NS ? ch : 0];
                    else { const int j = ch - S::NS >= 0 ? ch - S::NS : 0; val = v[j % 3][(j / 3) < S::NV ? j / 3 : 0]; }
                    M[c * 129 + row] = val;
                }
            }
            wg_sync(wg);
            const int span = n_last - n_first + 1;
            const int nch = min(S::CHH, S::CH - ch0);
            for (int i = row; i < span * nch; i += 128) {
                const int n = n_first + i / nch, c = i % nch, ch = ch0 + c;
                const long long ra_ = __ldg(a.rowptr + n), rb_ = __ldg(a.rowptr + n + 1);
                const int ra = (int)(max(ra_, p0) - p0), rb = (int)(min(rb_, p1) - p0);
                if (ra >= rb) continue;
                float sum = 0.f;
#pragma unroll 4
                for (int r = ra; r < rb; ++r) sum += M[c * 129 + r];
                if (ra_ >= p0 && rb_ <= p1) {
                    const float f = a.mean ? 1.f / (float)max((int)(rb_ - ra_), 1) : 1.f;
                    if (ch < S::NS) a.out_s[(long long)n * S::NS + ch] = sum * f;
                    else a.out_v[(long long)n * 3 * S::NV + (ch - S::NS)] = sum * f;
                } else if (ra_ < p0) {
                    a.part_head[(long long)t * S::CH + ch] = sum;
                } else {
                    a.part_tail[(long long)t * S::CH + ch] = sum;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(*slot, 512);
}

// ---- host side --------------------------------------------------------------------------------------------------------------
using TcMb = TcSpec<100, 16, 32, 1>;

static bool g_tensor_cores = false;
extern "C" int32_t cgvp_set_tensor_cores(int32_t on) { g_tensor_cores = on != 0; return 0; }

__global__ void conv_fixup_kernel(long long N, int CH, int SW, const int* __restrict__ rowptr, int mean, int tile_shift,
                                  const float* __restrict__ part_head, const float* __restrict__ part_tail,
                                  float* __restrict__ out_s, float* __restrict__ out_v);

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int64_t conv_tc_workspace_bytes(const CgvpConvDesc* desc, int64_t E, int64_t N) {
    using S = TcMb;
    if (!S::matches(*desc)) return 0;
    int64_t b = align_up(S::W_BYTES, 256) + align_up(S::WHE_FLOATS * 4, 256);
    b += 2 * align_up(N * S::NS * 4, 256) + 2 * align_up(N * 3 * S::HQ * 4, 256);
    b += 2 * align_up(cdiv64(E > 0 ? E : 1, 128) * S::CH * 4, 256);
    return b + 256;
}

// Returns 1 if the tensor-core kernel served the call (*rc_out = result), 0 otherwise.
int conv_fwd_tc(const CgvpConvDesc* desc, const CgvpPlan* plan, const float* x_s, const float* x_v, const float* e_s,
                const float* e_v, const float* const* h_packed, float* out_s, float* out_v, void* tcws, int64_t tcws_bytes,
                cudaStream_t st, int* rc_out) {
    using S = TcMb;
    if (!g_tensor_cores || !S::matches(*desc) || plan->num_edges <= 0 || plan->num_nodes <= 0) return 0;
    const int64_t E = plan->num_edges, N = plan->num_nodes;
    if (!tcws || tcws_bytes < conv_tc_workspace_bytes(desc, E, N) || !aligned16(e_s) || !aligned16(x_s) || !aligned16(x_v)) return 0;
    *rc_out = 0;
    char* b = reinterpret_cast<char*>(tcws);
    b = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(b) + 255) & ~(uintptr_t)255);
    unsigned char* wtc = reinterpret_cast<unsigned char*>(b); b += align_up(S::W_BYTES, 256);
    float* whe = reinterpret_cast<float*>(b); b += align_up(S::WHE_FLOATS * 4, 256);
    float* psj = reinterpret_cast<float*>(b); b += align_up(N * S::NS * 4, 256);
    float* psi = reinterpret_cast<float*>(b); b += align_up(N * S::NS * 4, 256);
    float* pvj = reinterpret_cast<float*>(b); b += align_up(N * 3 * S::HQ * 4, 256);
    float* pvi = reinterpret_cast<float*>(b); b += align_up(N * 3 * S::HQ * 4, 256);
    const int64_t ntiles = cdiv64(E, 128);
    float* part_head = reinterpret_cast<float*>(b); b += align_up(ntiles * S::CH * 4, 256);
    float* part_tail = reinterpret_cast<float*>(b);
    const int sms = cgvp_num_sms();
    auto fail = [&](cudaError_t e, const char* what) { cgvp_set_error("%s failed: %s", what, cudaGetErrorString(e)); *rc_out = (int)e; return 1; };
    tc_pack_kernel<S><<<cdiv(S::W_BYTES / 2, 256), 256, 0, st>>>(h_packed[0], h_packed[1], h_packed[2],
                                                                 reinterpret_cast<__nv_bfloat16*>(wtc), whe);
    tc_node_proj_kernel<S><<<(int)min((long long)cdiv64(N, 8), (long long)sms * 8), 256, 0, st>>>(N, x_s, x_v, h_packed[0], psj, psi, pvj, pvi);
    TcArgs a;
    memset(&a, 0, sizeof(a));
    a.E = E; a.N = N; a.ntiles = (int)ntiles; a.mean = desc->aggr == CGVP_AGGR_MEAN; a.edge_sorted = desc->edge_sorted;
    a.perm = plan->perm; a.src = plan->src; a.dst = plan->dst; a.rowptr = plan->rowptr;
    a.e_s = e_s; a.e_v = e_v; a.psj = psj; a.psi = psi; a.pvj = pvj; a.pvi = pvi; a.wtc = wtc; a.whe = whe;
    a.out_s = out_s; a.out_v = out_v; a.part_head = part_head; a.part_tail = part_tail;
    const size_t smem = S::smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(conv_tc_fwd_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(e, "cudaFuncSetAttribute(conv_tc_fwd_kernel)");
    const int grid = (int)min((long long)cdiv64(ntiles, 2), (long long)sms);
    cgvp_prof_begin(CGVP_K_CONV_FWD, st);
    conv_tc_fwd_kernel<S><<<grid, 256, smem, st>>>(a);
    cgvp_prof_end(CGVP_K_CONV_FWD, st);
    conv_fixup_kernel<<<(unsigned)cdiv64(N * S::CH, 256), 256, 0, st>>>(N, S::CH, S::NS, plan->rowptr, a.mean, 7, part_head, part_tail, out_s, out_v);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(e, "launch of conv_tc_fwd_kernel");
    return 1;
}

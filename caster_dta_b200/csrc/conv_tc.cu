// Tensor-core fused GVPConv forward (tcgen05 + TMEM, sm_100a) for wide node dims -- BASELINE config 5:
// nodes (100,16), edges (32,1), message chain GVP(relu,gate) -> GVP(relu,gate) -> GVP(none,gate).
// Replaces GVPConv.forward/message + PyG propagate (models/gvp_layers.py:291-308) like conv.cu / conv_reg.cu.
//
// Per CTA (one per SM, persistent, 512 threads): two groups of 256 threads, each owning one 128-edge tile at a time and
// ping-ponging on the tensor pipe.  Inside a group, edge row r (= TMEM lane r) is shared by TWO threads, each owning half
// of the row's columns (so 16 warps fit in the register file and the per-thread epilogue halves).  The projections
// of the three message GVPs are tcgen05.mma (kind::f16, bf16 operands, fp32 accumulation in TMEM, M = 128 edges):
//     A operand  = the tile's activations, written by the owning threads as bf16 into the canonical K-major
//                  no-swizzle layout  [k/8][row][8]  (8x16-byte core matrices, SBO = 128 B, LBO = 2048 B);
//     B operand  = weights, pre-packed once per call into the same layout and staged into shared memory with
//                  cp.async.bulk (TMA engine) on an mbarrier; they stay resident for the whole kernel;
//     D          = TMEM columns, read back with tcgen05.ld.32x32b for the fused epilogues: vector norms (:153),
//                  ReLU (:172), sigmoid gate (:158-163).
// The K / N orders of every GEMM are chosen so that each half-row thread reads and writes only contiguous chunks:
// output row of the scalar GEMM = [s'_h0 | gate_h0 | s'_h1 | gate_h1], its input row = [s_h0 | vn_h0 | s_h1 (+ ones) | vn_h1].
// Three algebraic folds keep the per-edge GEMM work and the number of MMA <-> epilogue round trips small:
//   (1) linearity before the gather: the s_j / s_i blocks of W_s and the V_j / V_i blocks of W_h of the FIRST message
//       GVP act on per-node rows, so they are applied once per node (tc_node_proj_kernel) and gathered; the per-edge
//       K of that GEMM drops from 265 to 66;
//   (2) the gate reads the PRE-activation s' (vector_act is None, :159-162), which is linear in [s ; vn ; 1]:
//       gate = (W_sv W_s) [s ; vn ; 1] + (W_sv b_s + b_g) rides the scalar GEMM as extra output columns;
//   (3) Vo = W_mu Vh = (W_mu W_h) V rides the W_h GEMM as extra output columns.
// Per tile that leaves 1 + 2 + 2 MMA batches.  Bias rows ride the GEMMs as a ones column.
// Aggregation: deterministic segmented sum over the sorted targets (per-tile pieces + conv_fixup_kernel).
// Accuracy: bf16 operands -> scale-relative error ~2e-3 (north star allows <= 1e-2 with tensor cores); the fp32
// paths (conv_reg.cu / conv.cu) remain the default -- see cgvp_set_tensor_cores().
#include <cuda_bf16.h>

#include "cgvp_reg.cuh"

using namespace cgvpr;

constexpr CGVP_HD inline int pad16(int x) { return (x + 15) / 16 * 16; }
constexpr CGVP_HD inline int pad8(int x) { return (x + 7) / 8 * 8; }

// where a fused column / K slot comes from
enum { SRC_ZERO = 0, SRC_S = 1, SRC_VN = 2, SRC_ONE = 3, SRC_ES = 4, SRC_GATE = 5, SRC_VH = 6, SRC_VO = 7 };
struct Slot { int kind, idx; };

template <int NS_, int NV_, int ES_, int EV_>
struct TcSpec {
    static constexpr int NS = NS_, NV = NV_, ES = ES_, EV = EV_, EV1 = max1(EV_);
    static constexpr int H0 = 2 * NV + EV;                       // hidden vector channels of message GVP 0
    static constexpr int SI0 = 2 * NS + ES, KSD0 = SI0 + H0, KSD1 = NS + NV;
    static constexpr int SOP = pad4(NS), VOP = pad4(NV), HP0 = pad4(H0), HP1 = pad4(NV);
    // per-half column groups
    static constexpr int N_S = pad16(NS), N_V = pad16(NV);
    static constexpr int SH = N_S / 2, GH = N_V / 2, HALF = SH + GH, N_SG = 2 * HALF, N_HV = 2 * N_V;
    static_assert(NS < N_S, "the ones column lives in a pad slot of s");
    static_assert(HALF % 16 == 0 && SH % 8 == 0 && GH % 8 == 0, "half rows are loaded 16 columns at a time");
    static_assert(ES % 16 == 0, "edge scalars are split evenly in 8-value chunks");
    static constexpr int VH0 = (H0 + 1) / 2, VH1 = H0 - VH0;     // stage-0 hidden vector channels per half
    static constexpr int VHA = pad4(VH0), VHB = pad4(VH1);       // their widths in the node tables
    static constexpr int VS = pad8(imax(VH0, VH1 + 1));          // A-operand slots per half (the ones column sits after VH1)
    static constexpr int ESH = ES / 2;
    static constexpr int K_S0 = 2 * (ESH + VS), K_H = 2 * GH, K_S1 = N_SG;
    static_assert(K_S0 % 16 == 0 && K_H % 16 == 0, "bf16 MMA K granularity");
    static constexpr int NSG = N_SG;                             // node-projected scalar row (same order as the GEMM output)
    static constexpr int PV_H1 = VHA + GH, PVW = VHA + GH + VHB + GH;   // node-projected vector row per plane
    // bf16 weight arena (bytes, [k/8][n][8] blocks)
    static constexpr int W_S0 = 0, W_ST1 = W_S0 + N_SG * K_S0 * 2;
    static constexpr int W_HV = 0, W_S = W_HV + N_HV * K_H * 2, W_STAGE = W_S + N_SG * K_S1 * 2;
    static constexpr int W_BYTES = W_ST1 + 2 * W_STAGE;
    // fp32 side arena (floats): edge-vector columns [EV][PVW], node-projection weights
    static constexpr int F_WHE = 0, F_WSN = pad4(EV1 * PVW), F_WVN = F_WSN + 2 * NS * NSG, F_END = F_WVN + 2 * NV * PVW;
    static constexpr int EF = F_WSN;                             // kept in shared memory
    // TMEM columns per group
    static constexpr int C_HV = 0, C_S = C_HV + 3 * N_HV, C_END = C_S + N_SG;
    static_assert(C_END <= 256, "TMEM columns per group");
    // activation tile region per group (bytes)
    static constexpr int A_S = 0, A_V = A_S + (K_S1 / 8) * 2048;
    static constexpr int CH = NS + 3 * NV, CHH = (CH + 1) / 2;   // message channels; reduced in two halves
    static constexpr int TILE_BYTES = imax(imax((K_S0 / 8) * 2048, A_V + 3 * (K_H / 8) * 2048), (int)align_up(CHH * 129 * 4, 16));
    static constexpr int GRP_BYTES = TILE_BYTES + 4 * 128 * 4 + 64;   // + src/dst/eid/spare + mbarrier
    static constexpr size_t smem_bytes() { return 1024 + (size_t)W_BYTES + EF * 4 + 2 * (size_t)GRP_BYTES + 64; }

    // fused scalar output column n (also the layout of the node-projected scalar rows)
    static constexpr CGVP_HD Slot out_s(int n) {
        const int h = n / HALF, r = n % HALF;
        if (r < SH) return h * SH + r < NS ? Slot{SRC_S, h * SH + r} : Slot{SRC_ZERO, 0};
        return h * GH + (r - SH) < NV ? Slot{SRC_GATE, h * GH + (r - SH)} : Slot{SRC_ZERO, 0};
    }
    // K slot k of the scalar GEMM of stages >= 1: [s_h | vn_h] per half, the ones column at s index NS
    static constexpr CGVP_HD Slot in_s1(int k) {
        const int h = k / HALF, r = k % HALF;
        if (r < SH) { const int i = h * SH + r; return i < NS ? Slot{SRC_S, i} : (i == NS ? Slot{SRC_ONE, 0} : Slot{SRC_ZERO, 0}); }
        return h * GH + (r - SH) < NV ? Slot{SRC_VN, h * GH + (r - SH)} : Slot{SRC_ZERO, 0};
    }
    // K slot k of the scalar GEMM of stage 0: [e_s_h | vn_h (+ ones in half 1)] per half
    static constexpr CGVP_HD Slot in_s0(int k) {
        const int h = k / (ESH + VS), r = k % (ESH + VS);
        if (r < ESH) return Slot{SRC_ES, h * ESH + r};
        const int s = r - ESH;
        if (h == 0) return s < VH0 ? Slot{SRC_VN, s} : Slot{SRC_ZERO, 0};
        return s < VH1 ? Slot{SRC_VN, VH0 + s} : (s == VH1 ? Slot{SRC_ONE, 0} : Slot{SRC_ZERO, 0});
    }
    // output column n of the vector GEMM of stages >= 1: [Vh_h | Vo_h] per half
    static constexpr CGVP_HD Slot out_hv(int n) {
        const int h = n / (2 * GH), r = n % (2 * GH);
        const int c = h * GH + (r < GH ? r : r - GH);
        return c < NV ? Slot{r < GH ? SRC_VH : SRC_VO, c} : Slot{SRC_ZERO, 0};
    }
    // column n of a node-projected vector row (stage 0): [Vh_h0 | Vo_h0 | Vh_h1 | Vo_h1]
    static constexpr CGVP_HD Slot out_pv(int n) {
        if (n < VHA) return n < VH0 ? Slot{SRC_VH, n} : Slot{SRC_ZERO, 0};
        if (n < PV_H1) return n - VHA < NV ? Slot{SRC_VO, n - VHA} : Slot{SRC_ZERO, 0};
        if (n < PV_H1 + VHB) return n - PV_H1 < VH1 ? Slot{SRC_VH, VH0 + (n - PV_H1)} : Slot{SRC_ZERO, 0};
        return GH + (n - PV_H1 - VHB) < NV ? Slot{SRC_VO, GH + (n - PV_H1 - VHB)} : Slot{SRC_ZERO, 0};
    }
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
    const float *psj, *psi, *pvj, *pvi;     // per-node projections [N][NSG], [N][NSG], [N][3][PVW], [N][3][PVW]
    const unsigned char* wtc;               // bf16 weight arena
    const float* wf;                        // fp32 side arena
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
__device__ __forceinline__ void grp_sync(int g) { asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory"); }
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
// Generic fp32 packed blocks (cgvp_common.cuh) -> bf16 [k/8][n][8] blocks of the fused / permuted GEMM shapes, plus the
// fp32 side arena (node-projection weights and the edge-vector columns).
template <class S>
struct TcW {
    using G0 = GvpC<S::SI0, S::H0, S::NS, S::NV, S::H0, 1, 0, 1>;
    using G1 = GvpC<S::NS, S::NV, S::NS, S::NV, S::NV, 1, 0, 1>;
    // weight from ws input row `row` (a row of ws_t; bias row = KSD) to the fused output slot `o`
    template <class G>
    static __device__ float sg(const float* w, int row, Slot o, bool bias_row) {
        if (o.kind == SRC_S) return w[G::O_WS_T + row * S::SOP + o.idx];
        if (o.kind != SRC_GATE) return 0.f;
        float acc = bias_row ? w[G::O_WSV_T + S::NS * S::VOP + o.idx] : 0.f;
        for (int j = 0; j < S::NS; ++j) acc += w[G::O_WS_T + row * S::SOP + j] * w[G::O_WSV_T + j * S::VOP + o.idx];
        return acc;
    }
    // weight from wh input channel row `crow` to the fused vector output slot `o` (Vh channel or Vo channel)
    template <class G, int H, int HP>
    static __device__ float hv(const float* w, int crow, Slot o) {
        if (o.kind == SRC_VH) return w[G::O_WH_T + crow * HP + o.idx];
        if (o.kind != SRC_VO) return 0.f;
        float acc = 0.f;
        for (int h = 0; h < H; ++h) acc += w[G::O_WH_T + crow * HP + h] * w[G::O_WV_T + h * S::VOP + o.idx];
        return acc;
    }
};

template <class S>
__global__ void tc_pack_kernel(const float* __restrict__ w0, const float* __restrict__ w1, const float* __restrict__ w2,
                               __nv_bfloat16* __restrict__ out, float* __restrict__ wf) {
    using W = TcW<S>;
    using G0 = typename W::G0; using G1 = typename W::G1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < S::F_END) {                                    // ---- fp32 side arena
        float v = 0.f;
        if (i < S::F_WSN) {
            const int c = i / S::PVW, n = i % S::PVW;
            if (c < S::EV) v = W::template hv<G0, S::H0, S::HP0>(w0, S::NV + c, S::out_pv(n));
        } else if (i < S::F_WVN) {
            const int j = i - S::F_WSN, side = j / (S::NS * S::NSG), r = j % (S::NS * S::NSG), k = r / S::NSG, n = r % S::NSG;
            v = W::template sg<G0>(w0, side ? S::NS + S::ES + k : k, S::out_s(n), false);
        } else {
            const int j = i - S::F_WVN, side = j / (S::NV * S::PVW), r = j % (S::NV * S::PVW), c = r / S::PVW, n = r % S::PVW;
            v = W::template hv<G0, S::H0, S::HP0>(w0, side ? S::NV + S::EV + c : c, S::out_pv(n));
        }
        wf[i] = v;
    }
    if (i >= S::W_BYTES / 2) return;                       // ---- bf16 arena
    const int byte = 2 * i;
    float v = 0.f;
    int n, k;
    auto blk = [&](int off, int npad) {                    // element (n, k) inside a [k/8][npad][8] block
        const int j = (byte - off) / 2;
        k = (j / (npad * 8)) * 8 + (j & 7);
        n = (j >> 3) % npad;
    };
    if (byte < S::W_ST1) {                                 // GVP 0 scalar GEMM, edge part
        blk(S::W_S0, S::N_SG);
        const Slot in = S::in_s0(k);
        if (in.kind == SRC_ES) v = W::template sg<G0>(w0, S::NS + in.idx, S::out_s(n), false);
        else if (in.kind == SRC_VN) v = W::template sg<G0>(w0, S::SI0 + in.idx, S::out_s(n), false);
        else if (in.kind == SRC_ONE) v = W::template sg<G0>(w0, S::KSD0, S::out_s(n), true);
    } else {
        const int st = (byte - S::W_ST1) / S::W_STAGE;
        const int base = S::W_ST1 + st * S::W_STAGE;
        const float* w = st == 0 ? w1 : w2;
        if (byte - base < S::W_S) {                        // [Vh | Vo] = [W_h ; W_mu W_h] V
            blk(base + S::W_HV, S::N_HV);
            if (k < S::NV) v = W::template hv<G1, S::NV, S::HP1>(w, k, S::out_hv(n));
        } else {                                           // [s' | gate] from [s ; vn ; 1]
            blk(base + S::W_S, S::N_SG);
            const Slot in = S::in_s1(k);
            if (in.kind == SRC_S) v = W::template sg<G1>(w, in.idx, S::out_s(n), false);
            else if (in.kind == SRC_VN) v = W::template sg<G1>(w, S::NS + in.idx, S::out_s(n), false);
            else if (in.kind == SRC_ONE) v = W::template sg<G1>(w, S::KSD1, S::out_s(n), true);
        }
    }
    out[i] = __float2bfloat16_rn(v);
}

// ---- per-node projections of message GVP 0 (fp32) -------------------------------------------------------------------------
template <class S>
__global__ void __launch_bounds__(256) tc_node_proj_kernel(long long N, const float* __restrict__ x_s, const float* __restrict__ x_v,
                                                            const float* __restrict__ wf, float* __restrict__ psj,
                                                            float* __restrict__ psi, float* __restrict__ pvj, float* __restrict__ pvi) {
    constexpr int NB = 8;                                  // nodes per pass
    __shared__ __align__(16) float xs[S::NS][NB];
    __shared__ __align__(16) float xv[3 * S::NV][NB];      // [p * NV + c][node]
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
        for (int q = t; q < 2 * S::NSG; q += blockDim.x) {          // scalar projections: (side, output)
            const int side = q / S::NSG, o = q % S::NSG;
            const float* w = wf + S::F_WSN + side * S::NS * S::NSG + o;
#pragma unroll
            for (int j = 0; j < NB; ++j) acc[j] = 0.f;
            for (int k = 0; k < S::NS; ++k) {
                const float wk = __ldg(w + k * S::NSG);
#pragma unroll
                for (int j = 0; j < NB; ++j) acc[j] = fmaf(xs[k][j], wk, acc[j]);
            }
            float* out = side ? psi : psj;
#pragma unroll
            for (int j = 0; j < NB; ++j)
                if (n0 + j < N) out[(n0 + j) * S::NSG + o] = acc[j];
        }
        for (int q = t; q < 2 * 3 * S::PVW; q += blockDim.x) {      // vector projections: (side, plane, output)
            const int side = q / (3 * S::PVW), r = q % (3 * S::PVW), p = r / S::PVW, o = r % S::PVW;
            const float* w = wf + S::F_WVN + side * S::NV * S::PVW + o;
#pragma unroll
            for (int j = 0; j < NB; ++j) acc[j] = 0.f;
            for (int c = 0; c < S::NV; ++c) {
                const float wc = __ldg(w + c * S::PVW);
#pragma unroll
                for (int j = 0; j < NB; ++j) acc[j] = fmaf(xv[p * S::NV + c][j], wc, acc[j]);
            }
            float* out = side ? pvi : pvj;
#pragma unroll
            for (int j = 0; j < NB; ++j)
                if (n0 + j < N) out[((n0 + j) * 3 + p) * S::PVW + o] = acc[j];
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

struct GrpCtx {
    unsigned char* tile;
    uint64_t* bar;
    uint32_t tm, tm_grp, w0s;
    int grp, row;
    bool leader;
    const float* ef;
};

// One MMA batch of a group: publish the A tiles, let the leader issue, wait for completion.
#define TC_BATCH_BEGIN()            \
    fence_proxy_async();            \
    tc_fence_before();              \
    grp_sync(g.grp);                \
    if (g.leader) {                 \
        tc_fence_after();
#define TC_BATCH_COMMIT()           \
        mma_commit(g.bar);          \
    }
#define TC_BATCH_WAIT()             \
    mbar_wait(g.bar, phase);        \
    phase ^= 1;                     \
    tc_fence_after();

// this half's [s'_h | gate_h] (HALF columns) from TMEM (+ optional addend already in sg), vo in registers -> outputs
template <class S, int H, bool RELU, bool ADD>
__device__ __forceinline__ void finish_stage(uint32_t tm, float (&sg)[S::HALF], float (&vo)[3][S::GH]) {
    if (ADD) {
        float d[S::HALF];
#pragma unroll
        for (int c = 0; c < S::HALF / 16; ++c) tmem_ld16(tm + S::C_S + H * S::HALF + 16 * c, d + 16 * c);
        tmem_ld_wait(d);
#pragma unroll
        for (int j = 0; j < S::HALF; ++j) sg[j] += d[j];
    } else {
#pragma unroll
        for (int c = 0; c < S::HALF / 16; ++c) tmem_ld16(tm + S::C_S + H * S::HALF + 16 * c, sg + 16 * c);
        tmem_ld_wait(sg);
    }
#pragma unroll
    for (int c = 0; c < S::GH; ++c) {
        const float gt = fast_sigmoid(sg[S::SH + c]);                                    // :158-163
#pragma unroll
        for (int p = 0; p < 3; ++p) vo[p][c] *= gt;
    }
    if (RELU) {
#pragma unroll
        for (int k = 0; k < S::SH; ++k) sg[k] = fmaxf(sg[k], 0.f);                       // :172-173
    }
}

// Everything one half-row thread does for one tile up to (and including) writing its message channels for the reduce.
template <class S, int H>
__device__ __forceinline__ void tile_messages(const TcArgs& a, const GrpCtx& g, uint32_t& phase, int src, int dst, long long eid,
                                              float (&sg)[S::HALF], float (&v)[3][S::GH]) {
    constexpr int VHW = H == 0 ? S::VHA : S::VHB;          // this half's Vh width in the node tables
    constexpr int VHN = H == 0 ? S::VH0 : S::VH1;          // ... and its real channel count
    constexpr int PVO = H == 0 ? 0 : S::PV_H1;             // offset of this half inside a node-projected vector row
    unsigned char* tile = g.tile;
    const int row = g.row;
    // ================= message GVP 0 =================
    {
        float vh[3][VHW];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const float* pj = a.pvj + ((long long)src * 3 + q) * S::PVW + PVO;
            const float* pi = a.pvi + ((long long)dst * 3 + q) * S::PVW + PVO;
            ld_row<VHW>(pj, vh[q]);
            add_row<VHW>(pi, vh[q]);
            ld_row<S::GH>(pj + VHW, v[q]);
            add_row<S::GH>(pi + VHW, v[q]);
        }
        if constexpr (S::EV > 0) {
#pragma unroll
            for (int c = 0; c < S::EV; ++c) {
                const float* we = g.ef + S::F_WHE + c * S::PVW + PVO;
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const float e = __ldg(a.e_v + (eid * S::EV + c) * 3 + q);
#pragma unroll
                    for (int o = 0; o < VHN; ++o) vh[q][o] = fmaf(e, we[o], vh[q][o]);
#pragma unroll
                    for (int o = 0; o < S::GH; ++o) v[q][o] = fmaf(e, we[VHW + o], v[q][o]);
                }
            }
        }
        float es[S::ESH];
        ld_row<S::ESH>(a.e_s + eid * S::ES + H * S::ESH, es);
        constexpr int C0 = H * (S::ESH + S::VS) / 8;       // first chunk of this half
#pragma unroll
        for (int c = 0; c < (S::ESH + S::VS) / 8; ++c) {
            float q8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int r = 8 * c + j;
                if (r < S::ESH) q8[j] = es[r < S::ESH ? r : 0];
                else {
                    const int s_ = r - S::ESH;
                    if (s_ < VHN) {
                        const int o = (s_ >= 0 && s_ < VHN) ? s_ : 0;
                        q8[j] = fast_sqrt(fmaxf(vh[0][o] * vh[0][o] + vh[1][o] * vh[1][o] + vh[2][o] * vh[2][o], CGVP_EPS));   // :153
                    } else q8[j] = (H == 1 && s_ == S::VH1) ? 1.f : 0.f;
                }
            }
            put8(tile + S::A_S, C0 + c, row, q8);
        }
        TC_BATCH_BEGIN()
            issue_gemm<S::N_SG, S::K_S0>(g.tm_grp + S::C_S, smem_u32(tile + S::A_S), g.w0s + S::W_S0);
        TC_BATCH_COMMIT()
        ld_row<S::HALF>(a.psj + (long long)src * S::NSG + H * S::HALF, sg);     // node-projected part of [s' | gate]
        add_row<S::HALF>(a.psi + (long long)dst * S::NSG + H * S::HALF, sg);
        TC_BATCH_WAIT()
        finish_stage<S, H, true, true>(g.tm, sg, v);
    }
    // ================= message GVPs 1 and 2 =================
#pragma unroll
    for (int st = 0; st < 2; ++st) {
        const uint32_t wst = g.w0s + S::W_ST1 + st * S::W_STAGE;
#pragma unroll
        for (int q = 0; q < 3; ++q) put8(tile + S::A_V + q * (S::K_H / 8) * 2048, H, row, v[q]);
        TC_BATCH_BEGIN()
#pragma unroll
            for (int q = 0; q < 3; ++q)
                issue_gemm<S::N_HV, S::K_H>(g.tm_grp + S::C_HV + q * S::N_HV, smem_u32(tile + S::A_V + q * (S::K_H / 8) * 2048), wst + S::W_HV);
        TC_BATCH_COMMIT()
        // meanwhile: this half's scalar slots of the next A operand (s, the ones column, zero padding)
#pragma unroll
        for (int c = 0; c < S::SH / 8; ++c) {
            float q8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i = H * S::SH + 8 * c + j;
                q8[j] = i < S::NS ? sg[8 * c + j] : (i == S::NS ? 1.f : 0.f);
            }
            put8(tile + S::A_S, H * (S::HALF / 8) + c, row, q8);
        }
        TC_BATCH_WAIT()
        {
            float hv[3][2 * S::GH];
#pragma unroll
            for (int q = 0; q < 3; ++q) tmem_ld16(g.tm + S::C_HV + q * S::N_HV + H * 2 * S::GH, hv[q]);
            tmem_ld_wait(hv[0]); tmem_ld_wait(hv[1]); tmem_ld_wait(hv[2]);
            float vn[S::GH];
#pragma unroll
            for (int o = 0; o < S::GH; ++o) {
                vn[o] = fast_sqrt(fmaxf(hv[0][o] * hv[0][o] + hv[1][o] * hv[1][o] + hv[2][o] * hv[2][o], CGVP_EPS));     // :153
#pragma unroll
                for (int q = 0; q < 3; ++q) v[q][o] = hv[q][S::GH + o];
            }
            put8(tile + S::A_S, H * (S::HALF / 8) + S::SH / 8, row, vn);
        }
        TC_BATCH_BEGIN()
            issue_gemm<S::N_SG, S::K_S1>(g.tm_grp + S::C_S, smem_u32(tile + S::A_S), wst + S::W_S);
        TC_BATCH_COMMIT()
        TC_BATCH_WAIT()
        if (st == 0) finish_stage<S, H, true, false>(g.tm, sg, v);
        else finish_stage<S, H, false, false>(g.tm, sg, v);
    }
}

// write this thread's message channels that fall into [ch0, ch0 + CHH) into M[ch - ch0][row]
template <class S, int H>
__device__ __forceinline__ void stage_channels(float* M, int row, int ch0, const float (&sg)[S::HALF], const float (&v)[3][S::GH]) {
#pragma unroll
    for (int j = 0; j < S::SH; ++j) {
        const int ch = H * S::SH + j;
        if (ch < S::NS && ch >= ch0 && ch < ch0 + S::CHH) M[(ch - ch0) * 129 + row] = sg[j];
    }
#pragma unroll
    for (int c = 0; c < S::GH; ++c)
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            const int cc = H * S::GH + c, ch = S::NS + 3 * cc + p;
            if (cc < S::NV && ch >= ch0 && ch < ch0 + S::CHH) M[(ch - ch0) * 129 + row] = v[p][c];
        }
}

template <class S>
__global__ void __launch_bounds__(512, 1) conv_tc_fwd_kernel(const __grid_constant__ TcArgs a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* wsm = smem;
    float* ef = reinterpret_cast<float*>(smem + S::W_BYTES);
    unsigned char* gbase = smem + S::W_BYTES + S::EF * 4;
    const int tid = threadIdx.x, grp = tid >> 8, u = tid & 255, row = u & 127, half = u >> 7, warp = tid >> 5;
    unsigned char* tile = gbase + grp * S::GRP_BYTES;
    int* idst = reinterpret_cast<int*>(tile + S::TILE_BYTES);
    uint64_t* bar = reinterpret_cast<uint64_t*>(tile + S::TILE_BYTES + 4 * 128 * 4);
    uint64_t* wbar = reinterpret_cast<uint64_t*>(gbase + 2 * S::GRP_BYTES);
    uint32_t* slot = reinterpret_cast<uint32_t*>(wbar + 1);

    if (tid == 0) {
        mbar_init(wbar, 1);
        mbar_init(reinterpret_cast<uint64_t*>(gbase + S::TILE_BYTES + 4 * 128 * 4), 1);
        mbar_init(reinterpret_cast<uint64_t*>(gbase + S::GRP_BYTES + S::TILE_BYTES + 4 * 128 * 4), 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(slot, 512);
    for (int i = tid; i < S::EF; i += blockDim.x) ef[i] = a.wf[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {                                        // weights: TMA bulk copies, resident for the whole kernel
        mbar_expect_tx(wbar, S::W_BYTES);
        constexpr int CHUNK = 16384;
        for (int off = 0; off < S::W_BYTES; off += CHUNK)
            bulk_g2s(wsm + off, a.wtc + off, (uint32_t)(S::W_BYTES - off < CHUNK ? S::W_BYTES - off : CHUNK), wbar);
    }
    GrpCtx g;
    g.tile = tile; g.bar = bar; g.grp = grp; g.row = row; g.leader = u == 0; g.ef = ef;
    g.tm_grp = *slot + (uint32_t)(grp * 256);                                   // MMA destination (lane 0)
    g.tm = g.tm_grp + ((uint32_t)((warp & 3) * 32) << 16);                       // this thread's lane quarter
    g.w0s = smem_u32(wsm);
    mbar_wait(wbar, 0);
    uint32_t phase = 0;

    for (int t = blockIdx.x * 2 + grp; t < a.ntiles; t += gridDim.x * 2) {
        const long long p0 = (long long)t * 128;
        const int rv = (int)min(128LL, a.E - p0);
        const long long p = row < rv ? p0 + row : p0;       // idle rows replay the first edge (never stored)
        const int src = __ldg(a.src + p), dst = __ldg(a.dst + p);
        const long long eid = a.edge_sorted ? p : (long long)__ldg(a.perm + p);
        grp_sync(grp);                                      // previous tile's reduce is done with the tile region / index array
        if (half == 0) idst[row] = dst;

        float sg[S::HALF], v[3][S::GH];
        if (half == 0) tile_messages<S, 0>(a, g, phase, src, dst, eid, sg, v);
        else tile_messages<S, 1>(a, g, phase, src, dst, eid, sg, v);

        // ================= aggregation: segmented sum over the sorted targets, two channel halves =================
        float* M = reinterpret_cast<float*>(tile);
        const int n_first = idst[0], n_last = idst[rv - 1];
        const long long p1 = p0 + rv;
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            grp_sync(grp);                                  // tile region free (MMAs done; previous pass consumed)
            const int ch0 = pass * S::CHH;
            if (half == 0) stage_channels<S, 0>(M, row, ch0, sg, v);
            else stage_channels<S, 1>(M, row, ch0, sg, v);
            grp_sync(grp);
            const int span = n_last - n_first + 1;
            const int nch = min(S::CHH, S::CH - ch0);
            for (int i = u; i < span * nch; i += 256) {
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
    int64_t b = align_up(S::W_BYTES, 256) + align_up(S::F_END * 4, 256);
    b += 2 * align_up(N * S::NSG * 4, 256) + 2 * align_up(N * 3 * S::PVW * 4, 256);
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
    float* wf = reinterpret_cast<float*>(b); b += align_up(S::F_END * 4, 256);
    float* psj = reinterpret_cast<float*>(b); b += align_up(N * S::NSG * 4, 256);
    float* psi = reinterpret_cast<float*>(b); b += align_up(N * S::NSG * 4, 256);
    float* pvj = reinterpret_cast<float*>(b); b += align_up(N * 3 * S::PVW * 4, 256);
    float* pvi = reinterpret_cast<float*>(b); b += align_up(N * 3 * S::PVW * 4, 256);
    const int64_t ntiles = cdiv64(E, 128);
    float* part_head = reinterpret_cast<float*>(b); b += align_up(ntiles * S::CH * 4, 256);
    float* part_tail = reinterpret_cast<float*>(b);
    const int sms = cgvp_num_sms();
    auto fail = [&](cudaError_t e, const char* what) { cgvp_set_error("%s failed: %s", what, cudaGetErrorString(e)); *rc_out = (int)e; return 1; };
    const int pack_threads = S::W_BYTES / 2 > S::F_END ? S::W_BYTES / 2 : S::F_END;
    tc_pack_kernel<S><<<cdiv(pack_threads, 128), 128, 0, st>>>(h_packed[0], h_packed[1], h_packed[2], reinterpret_cast<__nv_bfloat16*>(wtc), wf);
    tc_node_proj_kernel<S><<<(int)min((long long)cdiv64(N, 8), (long long)sms * 8), 256, 0, st>>>(N, x_s, x_v, wf, psj, psi, pvj, pvi);
    TcArgs a;
    memset(&a, 0, sizeof(a));
    a.E = E; a.N = N; a.ntiles = (int)ntiles; a.mean = desc->aggr == CGVP_AGGR_MEAN; a.edge_sorted = desc->edge_sorted;
    a.perm = plan->perm; a.src = plan->src; a.dst = plan->dst; a.rowptr = plan->rowptr;
    a.e_s = e_s; a.e_v = e_v; a.psj = psj; a.psi = psi; a.pvj = pvj; a.pvi = pvi; a.wtc = wtc; a.wf = wf;
    a.out_s = out_s; a.out_v = out_v; a.part_head = part_head; a.part_tail = part_tail;
    const size_t smem = S::smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(conv_tc_fwd_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(e, "cudaFuncSetAttribute(conv_tc_fwd_kernel)");
    const int grid = (int)min((long long)cdiv64(ntiles, 2), (long long)sms);
    cgvp_prof_begin(CGVP_K_CONV_FWD, st);
    conv_tc_fwd_kernel<S><<<grid, 512, smem, st>>>(a);
    cgvp_prof_end(CGVP_K_CONV_FWD, st);
    conv_fixup_kernel<<<(unsigned)cdiv64(N * S::CH, 256), 256, 0, st>>>(N, S::CH, S::NS, plan->rowptr, a.mean, 7, part_head, part_tail, out_s, out_v);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(e, "launch of conv_tc_fwd_kernel");
    return 1;
}

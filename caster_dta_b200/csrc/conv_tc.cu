// Tensor-core fused GVPConv forward (tcgen05 + TMEM, sm_100a) for wide node dims -- BASELINE config 5:
// nodes (100,16), edges (32,1), message chain GVP(relu,gate) -> GVP(relu,gate) -> GVP(none,gate).
// Replaces GVPConv.forward/message + PyG propagate (models/gvp_layers.py:291-308) like conv.cu / conv_reg.cu.
//
// Per CTA (one per SM, persistent): two warpgroups, each owning one 128-edge tile at a time (thread = edge = TMEM
// lane), ping-pong on the tensor pipe.  The projections of the three message GVPs are tcgen05.mma (kind::f16, bf16
// operands, fp32 accumulation in TMEM, M = 128 edges):
//     A operand  = the tile's activations, written by the owning threads as bf16 into the canonical K-major
//                  no-swizzle layout  [k/8][row][8]  (8x16-byte core matrices, SBO = 128 B, LBO = 2064 B: the +16
//                  keeps the cooperative row copies, whose lanes walk the k-chunks, free of bank conflicts);
//     B operand  = weights, pre-packed once per call into the same layout and staged into shared memory with
//                  cp.async.bulk (TMA engine) on an mbarrier; they stay resident for the whole kernel;
//     D          = TMEM columns, read back with tcgen05.ld.32x32b (one row per thread) for the fused epilogues:
//                  vector norms (:153), ReLU (:172), sigmoid gate (:158-163).
// Three algebraic folds keep the number of MMA <-> epilogue round trips small:
//   (1) linearity before the gather, on the TARGET side only: the s_i block of W_s and the V_i block of W_h of the FIRST
//       message GVP act on per-node rows; edges are sorted by target, so a warp reads ~1-2 distinct projected rows
//       (broadcast loads) and adds them in the epilogue.  The SOURCE side is different for every edge: it goes through
//       the tensor core from bf16 node rows (tc_node_prep_kernel) that the warp copies cooperatively, 16 bytes per lane
//       with consecutive lanes on consecutive chunks of a row, straight into the A operand -- no conversion, no
//       per-thread row gathers (those cost one L1 wavefront per lane per load and bounded the previous version);
//   (2) the gate reads the PRE-activation s' (vector_act is None, :159-162), which is linear in [s ; vn ; 1]:
//       gate = (W_sv W_s) [s ; vn ; 1] + (W_sv b_s + b_g) rides the scalar GEMM as 16 extra output columns;
//   (3) Vo = W_mu Vh = (W_mu W_h) V rides the W_h GEMM as 16 extra output columns.
// Per tile that leaves 2 + 2 + 2 MMA batches (was 2 + 3 + 3).  bias rows ride the GEMMs as a ones column.
// Aggregation: deterministic segmented sum over the sorted targets (per-tile pieces + conv_fixup_kernel).
// Accuracy: bf16 operands -> scale-relative error ~2e-3 (north star allows <= 1e-2 with tensor cores); the fp32
// paths (conv_reg.cu / conv.cu) remain the default -- see cgvp_set_tensor_cores().
#include "cgvp_reg.cuh"
#include <cuda_bf16.h>

using namespace cgvpr;

constexpr CGVP_HD inline int pad16(int x) { return (x + 15) / 16 * 16; }

template <int NS_, int NV_, int ES_, int EV_>
struct TcSpec {
    static constexpr int NS = NS_, NV = NV_, ES = ES_, EV = EV_, EV1 = max1(EV_);
    static constexpr int H0 = 2 * NV + EV;                       // hidden vector channels of message GVP 0
    static constexpr int HQ = pad4(H0);
    static constexpr int SI0 = 2 * NS + ES, KSD0 = SI0 + H0, KSD1 = NS + NV;
    static constexpr int SOP = pad4(NS), VOP = pad4(NV), HP0 = pad4(H0), HP1 = pad4(NV);
    // fused scalar GEMM output row: [s' (NS -> N_S) | gate pre-activation (NV -> N_V)]
    static constexpr int N_S = pad16(NS), N_V = pad16(NV), N_SG = N_S + N_V, N_HV = 2 * N_V;
    static constexpr int NSG = pad4(NS + NV);                    // node-projected scalar row: [s' part | gate part]
    static constexpr int PVW = HQ + N_V;                         // node-projected vector row per plane: [Vh | Vo]
    static constexpr int N_H0 = pad16(H0), N_HV0 = N_H0 + N_V;   // message GVP 0: [Vh (H0 -> N_H0) | Vo]
    // K of the scalar GEMM of message GVP 0: [s_j (NS -> KJ) ; e_s ; 0.. | vn (H0) ; 1 ; 0..]; the first part lives in the
    // scalar operand region, the second (written after the norms) in the vector operand region, free by then
    static constexpr int KJ = (NS + 7) / 8 * 8, K_A0 = pad16(KJ + ES), K_VN0 = pad16(H0 + 1), K_S0 = K_A0 + K_VN0;
    static constexpr int K_H = pad16(NV), K_S1 = pad16(NS + NV + 1);
    // bf16 node row: [s (KJ) | V plane 0 (K_H) | plane 1 | plane 2], XCH 16-byte chunks
    static constexpr int XROW = KJ + 3 * K_H, XCH = XROW / 8;
    static_assert(ES % 8 == 0 && K_H % 8 == 0, "16-byte chunks");
    // bf16 weight arena (bytes, [k/8][n][8] blocks)
    static constexpr int W_S0 = 0, W_V0 = W_S0 + N_SG * K_S0 * 2, W_ST1 = W_V0 + N_HV0 * K_H * 2;
    static constexpr int W_HV = 0, W_S = W_HV + N_HV * K_H * 2, W_STAGE = W_S + N_SG * K_S1 * 2;
    static constexpr int W_BYTES = W_ST1 + 2 * W_STAGE;
    // fp32 side arena (floats): edge-vector columns, target-side node projections
    static constexpr int F_WHE = 0, F_WVOE = F_WHE + pad4(EV1 * HQ), F_WSN = F_WVOE + pad4(EV1 * N_V);
    static constexpr int F_WVN = F_WSN + NS * NSG, F_END = F_WVN + NV * PVW;
    static constexpr int EF = F_WSN;                             // per-edge extras kept in shared memory
    // TMEM columns per warpgroup: [s' | gate] at 0; the [Vh | Vo] planes of GVP 0 alias it (read out before the scalar
    // GEMM is issued), those of GVPs 1-2 follow it
    static constexpr int C_S = 0, C_HV0 = 0, C_HV = N_SG, C_END = imax(3 * N_HV0, N_SG + 3 * N_HV);
    static_assert(C_END <= 256, "TMEM columns per warpgroup");
    // activation tile region per warpgroup (bytes); APITCH = byte pitch between k-chunks
    static constexpr int APITCH = 2048 + 16;
    static constexpr int A_S = 0, A_V = A_S + (imax(K_A0, K_S1) / 8) * APITCH;
    // message channels, reduced in two halves of CHH (quads never straddle the halves or the s / V boundary);
    // staged row-major with pitch RP floats (RP % 32 == 12: conflict-free 16-byte row stores)
    static constexpr int CH = NS + 3 * NV, CHH = (CH / 2 + 3) / 4 * 4, RP = CHH;
    static_assert(NS % 4 == 0 && CH % 4 == 0 && RP % 32 == 12, "aggregation staging");
    static constexpr int TILE_BYTES = imax(A_V + imax(3 * (K_H / 8), K_VN0 / 8) * APITCH, 128 * RP * 4);
    static constexpr int WG_BYTES = TILE_BYTES + 4 * 128 * 4 + 64;   // + src/dst/eid/spare + mbarrier
    static constexpr size_t smem_bytes() { return 1024 + (size_t)W_BYTES + EF * 4 + 2 * (size_t)WG_BYTES + 64; }
    static_assert(smem_bytes() <= 232448, "shared memory per CTA");
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
    const __nv_bfloat16* xb;                // bf16 node rows [N][XROW]
    const float *psi, *pvi;                 // target-side node projections [N][NSG], [N][3][PVW]
    const unsigned char* wtc;               // bf16 weight arena
    const float* wf;                        // fp32 side arena
    float *out_s, *out_v, *part_head, *part_tail;
};

#include "cgvp_tc.cuh"
using namespace tcx;

// D[128 x N] (+)= A[128 x K] . B[N x K]^T, K in steps of 16 (two 16-byte k-chunks per instruction)
template <int N, int K, int AP>
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, bool accumulate = false) {
    constexpr uint32_t id = idesc_bf16(N);
#pragma unroll
    for (int k = 0; k < K / 16; ++k)
        mma_bf16(tmem_d, smem_desc(a_addr + k * 2 * AP, AP), smem_desc(b_addr + k * 2 * (N * 16), N * 16), id, accumulate || k > 0);
}

// ---- weight pre-packing -------------------------------------------------------------------------------------------------
// Generic fp32 packed blocks (cgvp_common.cuh) -> bf16 [k/8][n][8] blocks of the padded / fused GEMM shapes, plus the
// fp32 side arena (node-projection weights and the edge-vector columns).
template <class S>
struct TcW {
    using G0 = GvpC<S::SI0, S::H0, S::NS, S::NV, S::H0, 1, 0, 1>;
    using G1 = GvpC<S::NS, S::NV, S::NS, S::NV, S::NV, 1, 0, 1>;
    // fused scalar weight of input row `row` (a row of ws_t, bias row = KSD) and fused output column n
    template <class G>
    static __device__ float sg(const float* w, int row, int n, bool bias_row) {
        if (n < S::NS) return w[G::O_WS_T + row * S::SOP + n];
        const int o = n - S::N_S;
        if (o < 0 || o >= S::NV) return 0.f;
        float acc = bias_row ? w[G::O_WSV_T + S::NS * S::VOP + o] : 0.f;
        for (int j = 0; j < S::NS; ++j) acc += w[G::O_WS_T + row * S::SOP + j] * w[G::O_WSV_T + j * S::VOP + o];
        return acc;
    }
    // fused vector weight of input channel row `crow` (a row of wh_t) and output column n: [Vh (width vw) | Vo]
    template <class G, int H, int HP>
    static __device__ float hv(const float* w, int crow, int n, int vw) {
        if (n < vw) return n < H ? w[G::O_WH_T + crow * HP + n] : 0.f;
        const int o = n - vw;
        if (o >= S::NV) return 0.f;
        float acc = 0.f;
        for (int h = 0; h < H; ++h) acc += w[G::O_WH_T + crow * HP + h] * w[G::O_WV_T + h * S::VOP + o];
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
        if (i < S::F_WVOE) {
            const int c = i / S::HQ, o = i % S::HQ;
            if (c < S::EV) v = W::template hv<G0, S::H0, S::HP0>(w0, S::NV + c, o, S::HQ);
        } else if (i < S::F_WSN) {
            const int j = i - S::F_WVOE, c = j / S::N_V, o = j % S::N_V;
            if (c < S::EV) v = W::template hv<G0, S::H0, S::HP0>(w0, S::NV + c, S::HQ + o, S::HQ);
        } else if (i < S::F_WVN) {                         // target-side scalar rows [k][NSG]
            const int j = i - S::F_WSN, k = j / S::NSG, n = j % S::NSG;
            if (n < S::NS + S::NV) v = W::template sg<G0>(w0, S::NS + S::ES + k, n < S::NS ? n : S::N_S + (n - S::NS), false);
        } else {                                           // target-side vector rows [c][PVW] = [Vh | Vo]
            const int j = i - S::F_WVN, c = j / S::PVW, n = j % S::PVW;
            v = W::template hv<G0, S::H0, S::HP0>(w0, S::NV + S::EV + c, n, S::HQ);
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
    if (byte < S::W_V0) {                                  // GVP 0 scalar GEMM: [s_j ; 0 | e_s ; 0 | vn ; 1 ; 0]
        blk(S::W_S0, S::N_SG);
        if (k < S::NS) v = W::template sg<G0>(w0, k, n, false);
        else if (k >= S::KJ && k < S::KJ + S::ES) v = W::template sg<G0>(w0, S::NS + (k - S::KJ), n, false);
        else if (k >= S::K_A0 && k < S::K_A0 + S::H0) v = W::template sg<G0>(w0, S::SI0 + (k - S::K_A0), n, false);
        else if (k == S::K_A0 + S::H0) v = W::template sg<G0>(w0, S::KSD0, n, true);
    } else if (byte < S::W_ST1) {                          // GVP 0 [Vh | Vo] from V_j
        blk(S::W_V0, S::N_HV0);
        if (k < S::NV) v = W::template hv<G0, S::H0, S::HP0>(w0, k, n, S::N_H0);
    } else {
        const int st = (byte - S::W_ST1) / S::W_STAGE;
        const int base = S::W_ST1 + st * S::W_STAGE;
        const float* w = st == 0 ? w1 : w2;
        if (byte - base < S::W_S) {                        // [Vh | Vo] = [W_h ; W_mu W_h] V
            blk(base + S::W_HV, S::N_HV);
            if (k < S::NV) v = W::template hv<G1, S::NV, S::HP1>(w, k, n, S::N_V);
        } else {                                           // [s' | gate] from [s ; vn ; 1]
            blk(base + S::W_S, S::N_SG);
            if (k <= S::KSD1) v = W::template sg<G1>(w, k, n, k == S::KSD1);
        }
    }
    out[i] = __float2bfloat16_rn(v);
}

// ---- per-node preparation ------------------------------------------------------------------------------------------------
//   xb[n]     = bf16 [x_s[n] ; 0 | x_V[n][:, 0] | x_V[n][:, 1] | x_V[n][:, 2]]       source-side A-operand rows
//   psi[n]    = Wsn^T x_s[n]              target-side part of [s' | gate] of message GVP 0 (fp32, width NSG)
//   pvi[n][p] = Wvn^T x_V[n][:, p]        target-side part of [Vh | Vo]                    (fp32, width PVW)
template <class S>
__global__ void __launch_bounds__(256) tc_node_prep_kernel(long long N, const float* __restrict__ x_s, const float* __restrict__ x_v,
                                                            const float* __restrict__ wf, __nv_bfloat16* __restrict__ xb,
                                                            float* __restrict__ psi, float* __restrict__ pvi) {
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
        for (int i = t; i < S::XROW * NB; i += blockDim.x) {
            const int nn = i / S::XROW, e = i % S::XROW;
            float val = 0.f;
            if (e < S::NS) val = xs[e][nn];
            else if (e >= S::KJ) { const int p = (e - S::KJ) / S::K_H, c = (e - S::KJ) % S::K_H; if (c < S::NV) val = xv[p * S::NV + c][nn]; }
            if (n0 + nn < N) xb[(n0 + nn) * S::XROW + e] = __float2bfloat16_rn(val);
        }
        float acc[NB];
        for (int o = t; o < S::NSG; o += blockDim.x) {              // scalar projection
            const float* w = wf + S::F_WSN + o;
#pragma unroll
            for (int j = 0; j < NB; ++j) acc[j] = 0.f;
            for (int k = 0; k < S::NS; ++k) {
                const float wk = __ldg(w + k * S::NSG);
#pragma unroll
                for (int j = 0; j < NB; ++j) acc[j] = fmaf(xs[k][j], wk, acc[j]);
            }
#pragma unroll
            for (int j = 0; j < NB; ++j)
                if (n0 + j < N) psi[(n0 + j) * S::NSG + o] = acc[j];
        }
        for (int q = t; q < 3 * S::PVW; q += blockDim.x) {          // vector projection: (plane, output)
            const int p = q / S::PVW, o = q % S::PVW;
            const float* w = wf + S::F_WVN + o;
#pragma unroll
            for (int j = 0; j < NB; ++j) acc[j] = 0.f;
            for (int c = 0; c < S::NV; ++c) {
                const float wc = __ldg(w + c * S::PVW);
#pragma unroll
                for (int j = 0; j < NB; ++j) acc[j] = fmaf(xv[p * S::NV + c][j], wc, acc[j]);
            }
#pragma unroll
            for (int j = 0; j < NB; ++j)
                if (n0 + j < N) pvi[((n0 + j) * 3 + p) * S::PVW + o] = acc[j];
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

// One MMA batch of a warpgroup: publish the A tiles, let the leader issue, wait for completion.
#define TC_BATCH_BEGIN()            \
    fence_proxy_async();            \
    tc_fence_before();              \
    wg_sync(wg);                    \
    if (leader) {                   \
        tc_fence_after();
#define TC_BATCH_COMMIT()           \
        mma_commit(bar);            \
    }
#define TC_BATCH_WAIT()             \
    mbar_wait(bar, phase);          \
    phase ^= 1;                     \
    tc_fence_after();

// [s' | gate] in TMEM (+ optional addend already in sg) and vo in registers -> (s_out, V_out)
template <class S, bool RELU, bool ADD>
__device__ __forceinline__ void finish_stage(uint32_t tm, float (&sg)[S::N_SG], float (&vo)[3][S::N_V]) {
    if (ADD) {
        constexpr int NB = S::N_SG / 16, B0 = (NB + 1) / 2;
        float d0[16 * B0], d1[16 * (NB - B0 > 0 ? NB - B0 : 1)];
#pragma unroll
        for (int c = 0; c < B0; ++c) tmem_ld16(tm + S::C_S + 16 * c, d0 + 16 * c);
#pragma unroll
        for (int c = B0; c < NB; ++c) tmem_ld16(tm + S::C_S + 16 * c, d1 + 16 * (c - B0));
        tmem_ld_wait(d0); tmem_ld_wait(d1);
#pragma unroll
        for (int j = 0; j < 16 * B0; ++j) sg[j] += d0[j];
#pragma unroll
        for (int j = 16 * B0; j < 16 * NB; ++j) sg[j] += d1[j - 16 * B0];
    } else {
#pragma unroll
        for (int c = 0; c < S::N_SG / 16; ++c) tmem_ld16(tm + S::C_S + 16 * c, sg + 16 * c);
        tmem_ld_wait(sg);
    }
#pragma unroll
    for (int c = 0; c < S::NV; ++c) {
        const float g = fast_sigmoid(sg[S::N_S + c]);                                   // :158-163
#pragma unroll
        for (int p = 0; p < 3; ++p) vo[p][c] *= g;
    }
    if (RELU) {
#pragma unroll
        for (int k = 0; k < S::NS; ++k) sg[k] = fmaxf(sg[k], 0.f);                       // :172-173
    }
}

template <class S>
__global__ void __launch_bounds__(256, 1) conv_tc_fwd_kernel(const __grid_constant__ TcArgs a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* wsm = smem;
    float* ef = reinterpret_cast<float*>(smem + S::W_BYTES);           // whe | wvoe
    unsigned char* wgbase = smem + S::W_BYTES + S::EF * 4;
    const int tid = threadIdx.x, wg = tid >> 7, row = tid & 127, warp = tid >> 5;
    unsigned char* tile = wgbase + wg * S::WG_BYTES;
    int* isrc = reinterpret_cast<int*>(tile + S::TILE_BYTES);
    int* idst = isrc + 128;
    int* ieid = idst + 128;
    uint64_t* bar = reinterpret_cast<uint64_t*>(tile + S::TILE_BYTES + 4 * 128 * 4);
    uint64_t* wbar = reinterpret_cast<uint64_t*>(wgbase + 2 * S::WG_BYTES);
    uint32_t* slot = reinterpret_cast<uint32_t*>(wbar + 1);

    if (tid == 0) {
        mbar_init(wbar, 1);
        mbar_init(reinterpret_cast<uint64_t*>(wgbase + S::TILE_BYTES + 4 * 128 * 4), 1);
        mbar_init(reinterpret_cast<uint64_t*>(wgbase + S::WG_BYTES + S::TILE_BYTES + 4 * 128 * 4), 1);
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
    const uint32_t tm_wg = *slot + (uint32_t)(wg * 256);                          // MMA destination (lane 0)
    const uint32_t tm = tm_wg + ((uint32_t)((warp & 3) * 32) << 16);               // this thread's lane
    mbar_wait(wbar, 0);
    uint32_t phase = 0;
    const bool leader = row == 0;
    const uint32_t w0s = smem_u32(wsm);
    const float* whe = ef + S::F_WHE;
    const float* wvoe = ef + S::F_WVOE;

    for (int t = blockIdx.x * 2 + wg; t < a.ntiles; t += gridDim.x * 2) {
        const long long p0 = (long long)t * 128;
        const int rv = (int)min(128LL, a.E - p0);
        const long long p = row < rv ? p0 + row : p0;       // idle rows replay the first edge (never stored)
        const int src = __ldg(a.src + p), dst = __ldg(a.dst + p);
        const long long eid = a.edge_sorted ? p : (long long)__ldg(a.perm + p);
        wg_sync(wg);                                        // previous tile's reduce is done with the tile region / index arrays
        isrc[row] = src; idst[row] = dst; ieid[row] = (int)eid;

        float sg[S::N_SG], v[3][S::N_V];
        // ================= message GVP 0 =================
        {
            // source rows -> A operand: the warp's 32 rows x XCH chunks, consecutive lanes on consecutive chunks of a row
            {
                const int lane = tid & 31, wrow0 = row & ~31;
                const uint4* xb4 = reinterpret_cast<const uint4*>(a.xb);
#pragma unroll
                for (int i = 0; i < S::XCH; ++i) {
                    const int m = i * 32 + lane, r = m / S::XCH, c = m - r * S::XCH;
                    const int sr = __shfl_sync(0xffffffffu, src, r);
                    const uint4 q = __ldg(xb4 + (long long)sr * S::XCH + c);
                    unsigned char* d = c < S::KJ / 8 ? tile + S::A_S + c * S::APITCH : tile + S::A_V + (c - S::KJ / 8) * S::APITCH;
                    *reinterpret_cast<uint4*>(d + (wrow0 + r) * 16) = q;
                }
            }
            // e_s -> chunks after s_j; the K padding up to K_A0 must be finite (its weights are zero)
            {
                float es[S::ES];
                ld_row<S::ES>(a.e_s + eid * S::ES, es);
#pragma unroll
                for (int c = 0; c < S::ES / 8; ++c) {
                    float q8[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) q8[j] = es[8 * c + j];
                    put8<S::APITCH>(tile + S::A_S, S::KJ / 8 + c, row, q8);
                }
#pragma unroll
                for (int c = (S::KJ + S::ES) / 8; c < S::K_A0 / 8; ++c)
                    *reinterpret_cast<uint4*>(tile + S::A_S + c * S::APITCH + row * 16) = make_uint4(0u, 0u, 0u, 0u);
            }
            TC_BATCH_BEGIN()
#pragma unroll
                for (int q = 0; q < 3; ++q)
                    issue_gemm<S::N_HV0, S::K_H, S::APITCH>(tm_wg + S::C_HV0 + q * S::N_HV0,
                                                            smem_u32(tile + S::A_V + q * (S::K_H / 8) * S::APITCH), w0s + S::W_V0);
            TC_BATCH_COMMIT()
            // while the tensor pipe works: target-side projections + edge-vector columns of [Vh | Vo]           :152,:156
            float vh[3][S::HQ];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const float* pi = a.pvi + ((long long)dst * 3 + q) * S::PVW;
                ld_row<S::HQ>(pi, vh[q]);
                ld_row<S::N_V>(pi + S::HQ, v[q]);
            }
            if constexpr (S::EV > 0) {
#pragma unroll
                for (int c = 0; c < S::EV; ++c) {
                    float e3[3];
#pragma unroll
                    for (int q = 0; q < 3; ++q) e3[q] = __ldg(a.e_v + (eid * S::EV + c) * 3 + q);
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
#pragma unroll
                        for (int o = 0; o < S::H0; ++o) vh[q][o] = fmaf(e3[q], whe[c * S::HQ + o], vh[q][o]);
#pragma unroll
                        for (int o = 0; o < S::NV; ++o) v[q][o] = fmaf(e3[q], wvoe[c * S::N_V + o], v[q][o]);
                    }
                }
            }
            TC_BATCH_WAIT()
            // + source-side part from TMEM, 32 columns at a time
#pragma unroll
            for (int q = 0; q < 3; ++q) {
#pragma unroll
                for (int c0 = 0; c0 < S::N_HV0; c0 += 32) {
                    float d[32];
                    tmem_ld16(tm + S::C_HV0 + q * S::N_HV0 + c0, d);
                    tmem_ld16(tm + S::C_HV0 + q * S::N_HV0 + c0 + 16, d + 16);
                    tmem_ld_wait(d);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int o = c0 + j;
                        if (o < S::H0) vh[q][o] += d[j];
                        else if (o >= S::N_H0 && o < S::N_H0 + S::NV) v[q][o - S::N_H0] += d[j];
                    }
                }
            }
            // [vn ; 1 ; 0..] -> vector operand region (its V_j tiles have been consumed)
#pragma unroll
            for (int c = 0; c < S::K_VN0 / 8; ++c) {
                float q8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int k = 8 * c + j;
                    if (k < S::H0) q8[j] = fast_sqrt(fmaxf(vh[0][k] * vh[0][k] + vh[1][k] * vh[1][k] + vh[2][k] * vh[2][k], CGVP_EPS));   // :153
                    else q8[j] = k == S::H0 ? 1.f : 0.f;
                }
                put8<S::APITCH>(tile + S::A_V, c, row, q8);
            }
            TC_BATCH_BEGIN()
                issue_gemm<S::N_SG, S::K_A0, S::APITCH>(tm_wg + S::C_S, smem_u32(tile + S::A_S), w0s + S::W_S0);
                issue_gemm<S::N_SG, S::K_VN0, S::APITCH>(tm_wg + S::C_S, smem_u32(tile + S::A_V), w0s + S::W_S0 + (S::K_A0 / 8) * (S::N_SG * 16), true);
            TC_BATCH_COMMIT()
            // meanwhile: the target-side part of [s' | gate]
            {
                const float* ps = a.psi + (long long)dst * S::NSG;
                ld_row<S::NS>(ps, sg);
#pragma unroll
                for (int k = S::NS; k < S::N_S; ++k) sg[k] = 0.f;
                ld_row<S::N_V>(ps + S::NS, sg + S::N_S);
            }
            TC_BATCH_WAIT()
            finish_stage<S, true, true>(tm, sg, v);
        }
        // ================= message GVPs 1 and 2 =================
#pragma unroll
        for (int st = 0; st < 2; ++st) {
            const uint32_t wst = w0s + S::W_ST1 + st * S::W_STAGE;
            // [Vh | Vo] = [W_h ; W_mu W_h] V                                                                       :152,:156
#pragma unroll
            for (int q = 0; q < 3; ++q)
#pragma unroll
                for (int c = 0; c < S::K_H / 8; ++c) {
                    float q8[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) { const int o = 8 * c + j; q8[j] = o < S::NV ? v[q][o < S::NV ? o : 0] : 0.f; }
                    put8<S::APITCH>(tile + S::A_V + q * (S::K_H / 8) * S::APITCH, c, row, q8);
                }
            TC_BATCH_BEGIN()
#pragma unroll
                for (int q = 0; q < 3; ++q)
                    issue_gemm<S::N_HV, S::K_H, S::APITCH>(tm_wg + S::C_HV + q * S::N_HV, smem_u32(tile + S::A_V + q * (S::K_H / 8) * S::APITCH), wst + S::W_HV);
            TC_BATCH_COMMIT()
            // meanwhile: the scalar part of the next A operand
#pragma unroll
            for (int c = 0; c < S::NS / 8; ++c) {
                float q8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) q8[j] = sg[8 * c + j];
                put8<S::APITCH>(tile + S::A_S, c, row, q8);
            }
            TC_BATCH_WAIT()
            float vn[S::N_V];
            {
                float vh[3][S::N_V];
#pragma unroll
                for (int q = 0; q < 3; ++q) {
#pragma unroll
                    for (int c = 0; c < S::N_V / 16; ++c) {
                        tmem_ld16(tm + S::C_HV + q * S::N_HV + 16 * c, vh[q] + 16 * c);
                        tmem_ld16(tm + S::C_HV + q * S::N_HV + S::N_V + 16 * c, v[q] + 16 * c);
                    }
                }
                tmem_ld_wait(vh[0]); tmem_ld_wait(vh[1]); tmem_ld_wait(vh[2]);
                tmem_ld_wait(v[0]); tmem_ld_wait(v[1]); tmem_ld_wait(v[2]);
#pragma unroll
                for (int o = 0; o < S::N_V; ++o)
                    vn[o] = fast_sqrt(fmaxf(vh[0][o] * vh[0][o] + vh[1][o] * vh[1][o] + vh[2][o] * vh[2][o], CGVP_EPS));     // :153
            }
            // rest of [s ; vn ; 1]
#pragma unroll
            for (int c = S::NS / 8; c < S::K_S1 / 8; ++c) {
                float q8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int k = 8 * c + j;
                    if (k < S::NS) q8[j] = sg[k < S::NS ? k : 0];
                    else if (k < S::NS + S::NV) q8[j] = vn[(k - S::NS >= 0 && k - S::NS < S::NV) ? k - S::NS : 0];
                    else q8[j] = k == S::NS + S::NV ? 1.f : 0.f;
                }
                put8<S::APITCH>(tile + S::A_S, c, row, q8);
            }
            TC_BATCH_BEGIN()
                issue_gemm<S::N_SG, S::K_S1, S::APITCH>(tm_wg + S::C_S, smem_u32(tile + S::A_S), wst + S::W_S);
            TC_BATCH_COMMIT()
            TC_BATCH_WAIT()
            if (st == 0) finish_stage<S, true, false>(tm, sg, v);
            else finish_stage<S, false, false>(tm, sg, v);
        }
        // ================= aggregation: segmented sum over the sorted targets, two channel halves =================
        tc_fence_before();
        float* M = reinterpret_cast<float*>(tile);
        const int n_first = idst[0], n_last = idst[rv - 1];
        const long long p1 = p0 + rv;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            wg_sync(wg);                                    // tile region free (MMAs done; previous half consumed)
            constexpr int QH = S::CHH / 4;
            const int ch0 = half * S::CHH;
#pragma unroll
            for (int q = 0; q < QH; ++q) {
                float4 val;
                float* vp = &val.x;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ch = ch0 + 4 * q + j;
                    if (ch < S::NS) vp[j] = sg[ch < S::NS ? ch : 0];
                    else if (ch < S::CH) { const int e = ch - S::NS >= 0 ? ch - S::NS : 0; vp[j] = v[e % 3][(e / 3) < S::NV ? e / 3 : 0]; }
                    else vp[j] = 0.f;
                }
                if (ch0 + 4 * q < S::CH) *reinterpret_cast<float4*>(M + row * S::RP + 4 * q) = val;
            }
            wg_sync(wg);
            const int span = n_last - n_first + 1;
            const int nq = (min(S::CHH, S::CH - ch0)) / 4;
            for (int i = row; i < span * nq; i += 128) {
                const int n = n_first + i / nq, q = i % nq, ch = ch0 + 4 * q;
                const long long ra_ = __ldg(a.rowptr + n), rb_ = __ldg(a.rowptr + n + 1);
                const int ra = (int)(max(ra_, p0) - p0), rb = (int)(min(rb_, p1) - p0);
                if (ra >= rb) continue;
                float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
                for (int r = ra; r < rb; ++r) {
                    const float4 m = *reinterpret_cast<const float4*>(M + r * S::RP + 4 * q);
                    sum.x += m.x; sum.y += m.y; sum.z += m.z; sum.w += m.w;
                }
                float4* o;
                if (ra_ >= p0 && rb_ <= p1) {
                    const float f = a.mean ? 1.f / (float)max((int)(rb_ - ra_), 1) : 1.f;
                    sum.x *= f; sum.y *= f; sum.z *= f; sum.w *= f;
                    o = reinterpret_cast<float4*>(ch < S::NS ? a.out_s + (long long)n * S::NS + ch : a.out_v + (long long)n * 3 * S::NV + (ch - S::NS));
                } else {
                    o = reinterpret_cast<float4*>((ra_ < p0 ? a.part_head : a.part_tail) + (long long)t * S::CH + ch);
                }
                *o = sum;
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
    b += align_up(N * S::XROW * 2, 256) + align_up(N * S::NSG * 4, 256) + align_up(N * 3 * S::PVW * 4, 256);
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
    if (!tcws || tcws_bytes < conv_tc_workspace_bytes(desc, E, N) || !aligned16(e_s) || !aligned16(x_s) || !aligned16(x_v) || !aligned16(out_s) || !aligned16(out_v)) return 0;
    *rc_out = 0;
    char* b = reinterpret_cast<char*>(tcws);
    b = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(b) + 255) & ~(uintptr_t)255);
    unsigned char* wtc = reinterpret_cast<unsigned char*>(b); b += align_up(S::W_BYTES, 256);
    float* wf = reinterpret_cast<float*>(b); b += align_up(S::F_END * 4, 256);
    __nv_bfloat16* xb = reinterpret_cast<__nv_bfloat16*>(b); b += align_up(N * S::XROW * 2, 256);
    float* psi = reinterpret_cast<float*>(b); b += align_up(N * S::NSG * 4, 256);
    float* pvi = reinterpret_cast<float*>(b); b += align_up(N * 3 * S::PVW * 4, 256);
    const int64_t ntiles = cdiv64(E, 128);
    float* part_head = reinterpret_cast<float*>(b); b += align_up(ntiles * S::CH * 4, 256);
    float* part_tail = reinterpret_cast<float*>(b);
    const int sms = cgvp_num_sms();
    auto fail = [&](cudaError_t e, const char* what) { cgvp_set_error("%s failed: %s", what, cudaGetErrorString(e)); *rc_out = (int)e; return 1; };
    const int pack_threads = S::W_BYTES / 2 > S::F_END ? S::W_BYTES / 2 : S::F_END;
    tc_pack_kernel<S><<<cdiv(pack_threads, 128), 128, 0, st>>>(h_packed[0], h_packed[1], h_packed[2], reinterpret_cast<__nv_bfloat16*>(wtc), wf);
    tc_node_prep_kernel<S><<<(int)min((long long)cdiv64(N, 8), (long long)sms * 8), 256, 0, st>>>(N, x_s, x_v, wf, xb, psi, pvi);
    TcArgs a;
    memset(&a, 0, sizeof(a));
    a.E = E; a.N = N; a.ntiles = (int)ntiles; a.mean = desc->aggr == CGVP_AGGR_MEAN; a.edge_sorted = desc->edge_sorted;
    a.perm = plan->perm; a.src = plan->src; a.dst = plan->dst; a.rowptr = plan->rowptr;
    a.e_s = e_s; a.e_v = e_v; a.xb = xb; a.psi = psi; a.pvi = pvi; a.wtc = wtc; a.wf = wf;
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

// Register-resident fused GVPConv for compile-time known dims (the CASTER-DTA checkpoint dims (16,4)/(32,1)).
// Replaces GVPConv.forward/message + PyG propagate (models/gvp_layers.py:291-308) like conv.cu, but:
//   * one thread = one dst-sorted edge, the whole message chain lives in registers (cgvp_reg.cuh);
//   * one warp = 32 consecutive edges and is fully independent (no CTA barrier in the tile loop), weights are
//     shared-memory resident and read at warp-uniform addresses;
//   * the per-target aggregation is a warp-local segmented sum; pieces of segments that straddle 32-edge tiles
//     are combined in tile order by a small fix-up kernel (deterministic, no atomics, no memset of the output);
//   * backward recomputes each GVP from its stage input, stages the operands of the weight-gradient GEMMs in a
//     warp-private shared-memory buffer and accumulates them in a warp-private arena.
#include "cgvp_warp.cuh"

using namespace cgvpr;

int cgvp_segment_reduce_split(const float* rows, int width, const int* rowptr, const int* index, int64_t N, int aggr,
                              int beta, float* out_a, int wa, float* out_b, int wb, cudaStream_t st);

#define CR_WARPS 8
#define CR_THREADS (CR_WARPS * 32)

struct ConvRegArgs {
    long long E, N;
    int ntiles, mean, edge_sorted, acc_edge;
    const int *perm, *src, *dst, *rowptr;
    const float *x_s, *x_v, *e_s, *e_v;
    const float* wp[3];
    float *out_s, *out_v;                 // forward output / backward: target-side node gradient
    float *part_head, *part_tail;
    const float *d_out_s, *d_out_v;
    float *d_e_s, *d_e_v, *dj;
    float* partial;                       // [gridDim.x][PF] weight-gradient partials
    float* stash;                         // [E][STASH] training stash per sorted edge (forward writes, backward reads) or NULL
    const float* P;                       // [N][2 SO] per-node projections of the message scalars: [W_sj s_n ; W_si s_n + b]
    float *Ri, *Rj;                       // backward, [N][SO]: ds' of message GVP 0 summed over a node's in- / out-edges
    int part_row0;                        // node_post: first row of `partial` it may write
};

template <int NS_, int NV_, int ES_, int EV_, class G0_, class G1_, class G2_>
struct ConvSpec {
    static constexpr int NS = NS_, NV = NV_, ES = ES_, EV = EV_;
    using G0 = G0_; using G1 = G1_; using G2 = G2_;
    static_assert(G0::SI == 2 * NS + ES && G0::VI == 2 * NV + EV, "message input dims");
    static_assert(G1::SI == G0::SO && G1::VI == G0::VO && G2::SI == G1::SO && G2::VI == G1::VO, "chain dims");
    static_assert(G2::SO == NS && G2::VO == NV, "specialised conv maps node dims to node dims");
    static_assert(G0::SO == NS, "the node-level split of W_s reuses the node-row channel plan (ds' rides the scalar slots)");
    static constexpr int SO = G2::SO, VO = G2::VO;
    static constexpr int CH = SO + 3 * VO;        // message channels = output node row
    static constexpr int CHX = NS + 3 * NV;       // node row (gradient slices)
    // training stash row per sorted edge: [s'_0 ; V_1 ; s'_1 ; V_2 ; s'_2] -- the PRE-activation scalars of the three message
    // GVPs (the stage inputs follow by the activation) and the vector stage inputs: the backward skips every W_s projection
    static constexpr int ST1 = G0::SO + 3 * G0::VO, ST2 = G1::SO + 3 * G1::VO, STASH = ST1 + ST2 + G2::SO;
    // shared-memory weight offsets (floats)
    static constexpr int WF0 = 0, WF1 = G0::FWD_FLOATS, WF2 = WF1 + G1::FWD_FLOATS, WF = WF2 + G2::FWD_FLOATS;
    static constexpr int WT0 = 0, WT1 = G0::TOTAL_FLOATS, WT2 = WT1 + G1::TOTAL_FLOATS, WT = WT2 + G2::TOTAL_FLOATS;
    // gradient arena (= concatenated forward packed blocks)
    static constexpr int GO0 = 0, GO1 = G0::FWD_FLOATS, GO2 = GO1 + G1::FWD_FLOATS, PF = GO2 + G2::FWD_FLOATS;
    static constexpr int STG_COLS = imax(imax(sink_cols<G0>(), sink_cols<G1>()), sink_cols<G2>());
    static constexpr int STG_FLOATS = imax(STG_COLS * CGVP_WPITCH * 4, pad4(imax(CH, CHX) * CGVP_WPITCH));
    static constexpr size_t smem_fwd(bool const_weights) {
        return (const_weights ? 0 : (size_t)WF * 4) + (size_t)CR_WARPS * (CH * CGVP_WPITCH + 32) * 4;
    }
    static constexpr size_t smem_bwd(bool const_weights = false) {
        return (const_weights ? 0 : (size_t)WT * 4) + (size_t)CR_WARPS * (PF + STG_FLOATS + 32) * 4;
    }
    static bool matches(const CgvpConvDesc& d) {
        return d.ns == NS && d.nv == NV && d.es == ES && d.ev == EV && d.n_gvp == 3 && G0::matches(d.gvp[0]) &&
               G1::matches(d.gvp[1]) && G2::matches(d.gvp[2]);
    }
};

__device__ __forceinline__ void copy_f4w(float* dst, const float* __restrict__ src, int nfloats) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = threadIdx.x; i < (nfloats >> 2); i += blockDim.x) d4[i] = __ldg(s4 + i);
}

// message input (s_j, e_s, s_i), (V_j, e_V, V_i)  -- gvp_layers.py:306
template <class S>
__device__ __forceinline__ void load_message_input(const ConvRegArgs& a, int src, int dst, long long eid,
                                                   float (&s0)[1][S::G0::SI], float (&v0)[3][S::G0::VI1]) {
    load_s<S::NS, 0>(a.x_s, src, s0);
    load_s<S::ES, S::NS>(a.e_s, eid, s0);
    load_s<S::NS, S::NS + S::ES>(a.x_s, dst, s0);
    load_v<S::NV, 0>(a.x_v, src, v0);
    load_v<S::EV, S::NV>(a.e_v, eid, v0);
    load_v<S::NV, S::NV + S::EV>(a.x_v, dst, v0);
}

// Forward weights of the three message GVPs in the constant bank (cudaMemcpyToSymbolAsync before the launch, stream ordered):
// the weight operands then arrive through LDCU.128 in UNIFORM registers and feed FFMA2 directly -- no shared-memory crossbar
// traffic (a warp-uniform LDS.128 per two FFMA2 saturates the 128 B/clk crossbar at half the FMA rate; measured with
// scripts/microbench/weights_src.cu: 37 -> 49 TFLOP/s) and no vector registers for weights, which is what lets three CTAs
// share an SM.  One slot per device: calls that use it must be ordered on one stream (the encoder's main stream).
__constant__ float c_conv_fw[4096];
__constant__ float c_conv_bw[5120];          // forward + data-gradient halves for the backward kernel

// WSRC: 0 = weights in shared memory (LDS.128 broadcast), 1 = weights in the constant bank (c_conv_fw)
// The same with the node scalars replaced by their per-node projections (`a.P`): s0 receives the edge scalars only (slice
// [NS, NS + ES)), sp the start value of s' = P_j[src] + P_i[dst] (bias included), to be completed by gvp_fwd<G0, 2, NS, ES>.
template <class S>
__device__ __forceinline__ void load_message_edge(const ConvRegArgs& a, int src, int dst, long long eid,
                                                  float (&s0)[1][S::G0::SI], float (&v0)[3][S::G0::VI1],
                                                  float (&sp)[1][S::G0::SO]) {
    float pj[1][S::G0::SO], pi[1][S::G0::SO];
    load_s<S::G0::SO, 0>(a.P, 2LL * src, pj);
    load_s<S::G0::SO, 0>(a.P, 2LL * dst + 1, pi);
#pragma unroll
    for (int c = 0; c < S::G0::SO; ++c) sp[0][c] = pj[0][c] + pi[0][c];
#pragma unroll
    for (int c = 0; c < S::G0::SI; ++c) s0[0][c] = 0.f;            // the node slices are never read (dead code)
    load_s<S::ES, S::NS>(a.e_s, eid, s0);
    load_v<S::NV, 0>(a.x_v, src, v0);
    load_v<S::EV, S::NV>(a.e_v, eid, v0);
    load_v<S::NV, S::NV + S::EV>(a.x_v, dst, v0);
}

// P[n] = [W_sj s_n ; W_si s_n + b]: the rows of message GVP 0's W_s that multiply the source / target node scalars, applied
// once per node (the message input is [s_j ; e_s ; s_i], gvp_layers.py:306, so W_s [s_j ; e_s ; s_i ; vn] + b splits).
template <class S>
__global__ void __launch_bounds__(128) conv_node_proj_kernel(const __grid_constant__ ConvRegArgs a) {
    using G0 = typename S::G0;
    __shared__ __align__(16) float w[2 * S::NS * G0::SOP + G0::SOP];
    const float* ws = a.wp[0] + G0::O_WS_T;
    for (int i = threadIdx.x; i < S::NS * G0::SOP; i += blockDim.x) {
        w[i] = ws[i];                                                      // rows of s_j
        w[S::NS * G0::SOP + i] = ws[(S::NS + S::ES) * G0::SOP + i];        // rows of s_i
    }
    for (int i = threadIdx.x; i < G0::SOP; i += blockDim.x) w[2 * S::NS * G0::SOP + i] = ws[G0::KSD * G0::SOP + i];   // bias row
    __syncthreads();
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= a.N) return;
    float s[1][S::NS], pj[1][G0::SO], pi[1][G0::SO];
    load_s<S::NS, 0>(a.x_s, n, s);
#pragma unroll
    for (int o = 0; o < G0::SO; ++o) { pj[0][o] = 0.f; pi[0][o] = w[2 * S::NS * G0::SOP + o]; }
    mv<S::NS, G0::SO, G0::SOP, 0, 0>(w, s, pj);
    mv<S::NS, G0::SO, G0::SOP, 0, 0>(w + S::NS * G0::SOP, s, pi);
    float* P = const_cast<float*>(a.P);
    store_s<G0::SO, 0>(P, 2 * n, pj, false);
    store_s<G0::SO, 0>(P, 2 * n + 1, pi, false);
}

// Backward partner: Ri[n] / Rj[n] = ds' of message GVP 0 summed over the in- / out-edges of node n.
//   d_x_s[n]  = W_si^T Ri[n] + W_sj^T Rj[n]
//   dW_s rows of s_j += s_n (x) Rj[n],  rows of s_i += s_n (x) Ri[n],  bias row += Ri[n]     (every edge has one target)
// One thread per node, the weight gradients through the warp sink; each CTA leaves one row of `partial` (zero outside the
// W_s block of GVP 0), which the fixed-order reduction of the edge kernel's partials picks up.
template <class S>
__global__ void __launch_bounds__(128) conv_node_post_kernel(const __grid_constant__ ConvRegArgs a) {
    using G0 = typename S::G0;
    constexpr int NW = 4, WSB = pad4(G0::KS) * G0::SOP;        // the packed W_s block (rows 0 .. KSD = bias, padded to 4 rows)
    constexpr int ARENA = WSB + 4 * G0::SOP;                   // + the 4-row block the single bias row is added through
    constexpr int STG = 8 * CGVP_WPITCH * 4;
    __shared__ __align__(16) float wb[2 * G0::SOP * S::NS];    // ws_b columns of s_j, then of s_i: [o][k]
    __shared__ __align__(16) float arena0[NW * (ARENA + STG)];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* arena = arena0 + warp * (ARENA + STG);
    const float* wsb = a.wp[0] + G0::O_WS_B;
    for (int i = threadIdx.x; i < G0::SO * S::NS; i += blockDim.x) {
        const int o = i / S::NS, k = i - o * S::NS;
        wb[i] = wsb[o * G0::KSDP + k];
        wb[G0::SOP * S::NS + i] = wsb[o * G0::KSDP + S::NS + S::ES + k];
    }
    for (int i = lane; i < ARENA; i += 32) arena[i] = 0.f;
    __syncthreads();
    const float one[1][1] = {{1.f}};
    for (long long base = (long long)blockIdx.x * blockDim.x; base < a.N; base += (long long)gridDim.x * blockDim.x) {
        const long long n_ = base + threadIdx.x;
        const bool valid = n_ < a.N;
        const long long n = valid ? n_ : 0;
        WarpSink sink{reinterpret_cast<float4*>(arena + ARENA), arena, lane, valid};
        float s[1][S::NS], ri[1][G0::SO], rj[1][G0::SO], dx[1][S::NS];
        load_s<S::NS, 0>(a.x_s, n, s);
        load_s<G0::SO, 0>(a.Ri, n, ri);
        load_s<G0::SO, 0>(a.Rj, n, rj);
#pragma unroll
        for (int k = 0; k < S::NS; ++k) dx[0][k] = 0.f;
        mv<G0::SO, S::NS, S::NS, 0, 0>(wb, rj, dx);
        mv<G0::SO, S::NS, S::NS, 0, 0>(wb + G0::SOP * S::NS, ri, dx);
        if (valid) store_s<S::NS, 0>(a.out_s, n, dx, false);
        sink.template add<S::NS, G0::SO, 1>(0, s, rj);                                   // rows of s_j
        sink.template add<S::NS, G0::SO, 1>((S::NS + S::ES) * G0::SOP, s, ri);           // rows of s_i
        sink.template add<1, G0::SO, 1>(G0::KSD * G0::SOP, one, ri);                     // bias row
    }
    __syncthreads();
    float* out = a.partial + (long long)(a.part_row0 + blockIdx.x) * S::PF;
    for (int i = threadIdx.x; i < S::PF; i += blockDim.x) {
        float sum = 0.f;
        const int j = i - (S::GO0 + G0::O_WS_T);
        if (j >= 0 && j < WSB) {
#pragma unroll
            for (int w = 0; w < NW; ++w) sum += arena0[w * (ARENA + STG) + j];
        }
        out[i] = sum;
    }
}

template <class S, int WSRC, int MINB>
__global__ void __launch_bounds__(CR_THREADS, MINB) conv_fwd_reg_kernel(const __grid_constant__ ConvRegArgs a) {
    using G0 = typename S::G0; using G1 = typename S::G1; using G2 = typename S::G2;
    static_assert(S::WF <= 4096, "constant weight slot too small");
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* wsm;
    float* M;
    if constexpr (WSRC == 0) {
        float* w = reinterpret_cast<float*>(smem);
        copy_f4w(w + S::WF0, a.wp[0], G0::FWD_FLOATS);
        copy_f4w(w + S::WF1, a.wp[1], G1::FWD_FLOATS);
        copy_f4w(w + S::WF2, a.wp[2], G2::FWD_FLOATS);
        __syncthreads();
        wsm = w;
        M = w + S::WF + warp * (S::CH * CGVP_WPITCH + 32);
    } else {
        wsm = c_conv_fw;
        M = reinterpret_cast<float*>(smem) + warp * (S::CH * CGVP_WPITCH + 32);
    }
    int* dsts = reinterpret_cast<int*>(M + S::CH * CGVP_WPITCH);
    const int nw = gridDim.x * CR_WARPS;
    for (int t = blockIdx.x * CR_WARPS + warp; t < a.ntiles; t += nw) {
        const long long p0 = (long long)t * 32;
        const int rv = (int)min(32LL, a.E - p0);
        const long long p = lane < rv ? p0 + lane : p0;          // idle lanes replay the first row (never stored)
        const int src = __ldg(a.src + p), dst = __ldg(a.dst + p);
        const long long eid = a.edge_sorted ? p : (long long)__ldg(a.perm + p);
        const bool stash = a.stash && lane < rv;                // training: the backward reads these instead of recomputing them
        float* srow = a.stash + p * S::STASH;
        float s1[1][G0::SO], v1[3][G0::VO1];
        {
            float s0[1][G0::SI], v0[3][G0::VI1];
            Save<G0> sv;
            load_message_edge<S>(a, src, dst, eid, s0, v0, sv.sp);
            gvp_fwd<G0, 2, S::NS, S::ES>(wsm + S::WF0, s0, v0, s1, v1, sv);
            if (stash) {
                store_s<G0::SO, 0>(srow, 0, sv.sp, false);
                store_v<G0::VO, 0>(srow + G0::SO, 0, v1, false);
            }
        }
        float s2[1][G1::SO], v2[3][G1::VO1];
        {
            Save<G1> sv;
            gvp_fwd<G1>(wsm + S::WF1, s1, v1, s2, v2, sv);
            if (stash) {
                store_s<G1::SO, 0>(srow + S::ST1, 0, sv.sp, false);
                store_v<G1::VO, 0>(srow + S::ST1 + G1::SO, 0, v2, false);
            }
        }
        float s3[1][G2::SO], v3[3][G2::VO1];
        {
            Save<G2> sv;
            gvp_fwd<G2>(wsm + S::WF2, s2, v2, s3, v3, sv);
            if (stash) store_s<G2::SO, 0>(srow + S::ST1 + S::ST2, 0, sv.sp, false);
        }
#pragma unroll
        for (int c = 0; c < S::SO; ++c) M[c * CGVP_WPITCH + lane] = s3[0][c];
#pragma unroll
        for (int c = 0; c < S::VO; ++c)
#pragma unroll
            for (int q = 0; q < 3; ++q) M[(S::SO + 3 * c + q) * CGVP_WPITCH + lane] = v3[q][c];
        dsts[lane] = dst;
        __syncwarp();
        seg_reduce_warp<S::CH, S::SO>(M, dsts, lane, rv, p0, t, a.rowptr, a.mean != 0, a.out_s, a.out_v, a.part_head, a.part_tail);
        __syncwarp();
    }
}

// Nodes whose segment straddles 32-edge tiles (pieces summed in tile order) and nodes without incoming edges (zero,
// PyG scatter with dim_size = N).  All other nodes were written by the main kernel.
__global__ void conv_fixup_kernel(long long N, int CH, int SW, const int* __restrict__ rowptr, int mean, int tile_shift,
                                  const float* __restrict__ part_head, const float* __restrict__ part_tail,
                                  float* __restrict__ out_s, float* __restrict__ out_v) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * CH) return;
    const long long n = i / CH;
    const int ch = (int)(i - n * CH);
    const long long a = rowptr[n], b = rowptr[n + 1];
    float val;
    if (b <= a) {
        val = 0.f;
    } else {
        const long long ta = a >> tile_shift, tb = (b - 1) >> tile_shift;
        if (ta == tb) return;
        float sum = part_tail[ta * CH + ch];
        for (long long t = ta + 1; t <= tb; ++t) sum += part_head[t * CH + ch];
        val = mean ? sum / (float)max((int)(b - a), 1) : sum;
    }
    if (ch < SW) out_s[n * SW + ch] = val;
    else out_v[n * (CH - SW) + (ch - SW)] = val;
}

template <class S, int WSRC>
__global__ void __launch_bounds__(CR_THREADS, 1) conv_bwd_reg_kernel(const __grid_constant__ ConvRegArgs a) {
    using G0 = typename S::G0; using G1 = typename S::G1; using G2 = typename S::G2;
    static_assert(S::WT <= 5120, "constant weight slot too small");
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* wsm;
    float* arena0;
    if constexpr (WSRC == 0) {
        float* w = reinterpret_cast<float*>(smem);
        copy_f4w(w + S::WT0, a.wp[0], G0::TOTAL_FLOATS);
        copy_f4w(w + S::WT1, a.wp[1], G1::TOTAL_FLOATS);
        copy_f4w(w + S::WT2, a.wp[2], G2::TOTAL_FLOATS);
        wsm = w;
        arena0 = w + S::WT;
    } else {
        wsm = c_conv_bw;
        arena0 = reinterpret_cast<float*>(smem);
    }
    float* arena = arena0 + warp * (S::PF + S::STG_FLOATS + 32);
    float* stgf = arena + S::PF;
    int* dsts = reinterpret_cast<int*>(stgf + S::STG_FLOATS);
    for (int i = lane; i < S::PF; i += 32) arena[i] = 0.f;
    __syncthreads();
    const int nw = gridDim.x * CR_WARPS;
    for (int t = blockIdx.x * CR_WARPS + warp; t < a.ntiles; t += nw) {
        const long long p0 = (long long)t * 32;
        const int rv = (int)min(32LL, a.E - p0);
        const bool valid = lane < rv;
        const long long p = valid ? p0 + lane : p0;
        const int src = __ldg(a.src + p), dst = __ldg(a.dst + p);
        const long long eid = a.edge_sorted ? p : (long long)__ldg(a.perm + p);
        WarpSink sink{reinterpret_cast<float4*>(stgf), arena, lane, valid};
        // pre-activation scalars s'_k of the three message GVPs and the vector stage inputs: from the forward's training
        // stash, or from a forward chain.  Each GVP's remaining internals (Vh, norms, Vo, gate) are recomputed right before
        // its backward; its W_s projection is not (gvp_fwd<G, true>).
        float sp0[1][G0::SO], v1[3][G0::VO1], sp1[1][G1::SO], v2[3][G1::VO1], sp2[1][G2::SO];
        if (a.stash) {                                            // written by the forward pass of the same call pair
            const float* row = a.stash + p * S::STASH;
            load_s<G0::SO, 0>(row, 0, sp0);
            load_v<G0::VO, 0>(row + G0::SO, 0, v1);
            load_s<G1::SO, 0>(row + S::ST1, 0, sp1);
            load_v<G1::VO, 0>(row + S::ST1 + G1::SO, 0, v2);
            load_s<G2::SO, 0>(row + S::ST1 + S::ST2, 0, sp2);
        } else {
            float s1[1][G0::SO], s2[1][G1::SO];
            {
                float s0[1][G0::SI], v0[3][G0::VI1];
                Save<G0> sv;
                load_message_edge<S>(a, src, dst, eid, s0, v0, sv.sp);
                gvp_fwd<G0, 2, S::NS, S::ES>(wsm + S::WT0, s0, v0, s1, v1, sv);
#pragma unroll
                for (int c = 0; c < G0::SO; ++c) sp0[0][c] = sv.sp[0][c];
            }
            {
                Save<G1> sv;
                gvp_fwd<G1>(wsm + S::WT1, s1, v1, s2, v2, sv);
#pragma unroll
                for (int c = 0; c < G1::SO; ++c) sp1[0][c] = sv.sp[0][c];
            }
            {
                Save<G2> sv;
                float so[1][G2::SO], vo[3][G2::VO1];
                gvp_fwd<G2>(wsm + S::WT2, s2, v2, so, vo, sv);
#pragma unroll
                for (int c = 0; c < G2::SO; ++c) sp2[0][c] = sv.sp[0][c];
            }
        }
        // d(message_e) = d_out[dst_e] (/ deg for mean)
        float gs3[1][S::SO], gv3[3][G2::VO1];
        load_s<S::SO, 0>(a.d_out_s, dst, gs3);
        load_v<S::VO, 0>(a.d_out_v, dst, gv3);
        if (a.mean) {
            const float f = 1.f / (float)max(__ldg(a.rowptr + dst + 1) - __ldg(a.rowptr + dst), 1);
#pragma unroll
            for (int c = 0; c < S::SO; ++c) gs3[0][c] *= f;
#pragma unroll
            for (int q = 0; q < 3; ++q)
#pragma unroll
                for (int c = 0; c < S::VO; ++c) gv3[q][c] *= f;
        }
        float gs2[1][G1::SO], gv2[3][G1::VO1];
        {
            Save<G2> sv;
            float s2[1][G1::SO], so[1][G2::SO], vo[3][G2::VO1], dsin[1][G2::KSD], dvin[3][G2::VI1];
#pragma unroll
            for (int c = 0; c < G1::SO; ++c) s2[0][c] = actf<G1::SACT>(sp1[0][c]);
#pragma unroll
            for (int c = 0; c < G2::SO; ++c) sv.sp[0][c] = sp2[0][c];
            gvp_fwd<G2, 1>(wsm + S::WT2, s2, v2, so, vo, sv);
            gvp_bwd<G2>(wsm + S::WT2, sv, s2, v2, gs3, gv3, sink, S::GO2, dsin, dvin);
#pragma unroll
            for (int c = 0; c < G2::SI; ++c) gs2[0][c] = dsin[0][c];
#pragma unroll
            for (int q = 0; q < 3; ++q)
#pragma unroll
                for (int c = 0; c < G2::VI; ++c) gv2[q][c] = dvin[q][c];
        }
        float gs1[1][G0::SO], gv1[3][G0::VO1];
        {
            Save<G1> sv;
            float s1[1][G0::SO], so[1][G1::SO], vo[3][G1::VO1], dsin[1][G1::KSD], dvin[3][G1::VI1];
#pragma unroll
            for (int c = 0; c < G0::SO; ++c) s1[0][c] = actf<G0::SACT>(sp0[0][c]);
#pragma unroll
            for (int c = 0; c < G1::SO; ++c) sv.sp[0][c] = sp1[0][c];
            gvp_fwd<G1, 1>(wsm + S::WT1, s1, v1, so, vo, sv);
            gvp_bwd<G1>(wsm + S::WT1, sv, s1, v1, gs2, gv2, sink, S::GO1, dsin, dvin);
#pragma unroll
            for (int c = 0; c < G1::SI; ++c) gs1[0][c] = dsin[0][c];
#pragma unroll
            for (int q = 0; q < 3; ++q)
#pragma unroll
                for (int c = 0; c < G1::VI; ++c) gv1[q][c] = dvin[q][c];
        }
        // message GVP 0: only the edge scalars are handled per edge; the node scalars' share of W_s goes through ds'
        // (reduced per node -- Ri over the target, Rj over the source -- and finished by conv_node_post_kernel)
        float dsin[1][G0::KSD], dvin[3][G0::VI1], ds0[1][G0::SO];
        {
            float s0[1][G0::SI], v0[3][G0::VI1];
#pragma unroll
            for (int c = 0; c < G0::SI; ++c) s0[0][c] = 0.f;
            load_s<S::ES, S::NS>(a.e_s, eid, s0);
            load_v<S::NV, 0>(a.x_v, src, v0);
            load_v<S::EV, S::NV>(a.e_v, eid, v0);
            load_v<S::NV, S::NV + S::EV>(a.x_v, dst, v0);
            Save<G0> sv;
            float so[1][G0::SO], vo[3][G0::VO1];
#pragma unroll
            for (int c = 0; c < G0::SO; ++c) sv.sp[0][c] = sp0[0][c];
            gvp_fwd<G0, 1>(wsm + S::WT0, s0, v0, so, vo, sv);
            gvp_bwd_ds<G0, WarpSink, true, S::NS, S::ES>(wsm + S::WT0, sv, s0, v0, gs1, gv1, sink, S::GO0, dsin, dvin, ds0);
        }
        if (valid) {
            // edge-attribute gradient (one row per edge, written or accumulated)
            if (a.d_e_s) store_s<S::ES, S::NS>(a.d_e_s, eid, dsin, a.acc_edge != 0);
            if (a.d_e_v) store_v<S::EV, S::NV>(a.d_e_v, eid, dvin, a.acc_edge != 0);
            // source-side slice, one merged row per edge (reduced over the source CSR view afterwards)
            float dj[1][S::CHX];
#pragma unroll
            for (int c = 0; c < S::NS; ++c) dj[0][c] = ds0[0][c];
#pragma unroll
            for (int c = 0; c < S::NV; ++c)
#pragma unroll
                for (int q = 0; q < 3; ++q) dj[0][S::NS + 3 * c + q] = dvin[q][c];
            store_s<S::CHX, 0>(a.dj, p, dj, false);
        }
        // target-side slice: segmented sum over the target node
        float* M = stgf;
#pragma unroll
        for (int c = 0; c < S::NS; ++c) M[c * CGVP_WPITCH + lane] = ds0[0][c];
#pragma unroll
        for (int c = 0; c < S::NV; ++c)
#pragma unroll
            for (int q = 0; q < 3; ++q) M[(S::NS + 3 * c + q) * CGVP_WPITCH + lane] = dvin[q][S::NV + S::EV + c];
        dsts[lane] = dst;
        __syncwarp();
        seg_reduce_warp<S::CHX, S::NS>(M, dsts, lane, rv, p0, t, a.rowptr, false, a.out_s, a.out_v, a.part_head, a.part_tail);
        __syncwarp();
    }
    // CTA partial = sum of the warps' arenas in warp order (deterministic)
    __syncthreads();
    float* out = a.partial + (long long)blockIdx.x * S::PF;
    for (int i = threadIdx.x; i < S::PF; i += blockDim.x) {
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < CR_WARPS; ++w) sum += arena0[w * (S::PF + S::STG_FLOATS + 32) + i];
        out[i] = sum;
    }
}

// ---- instances --------------------------------------------------------------------------------------------------
// CASTER-DTA checkpoint dims: nodes (16,4), edges (32,1), message GVPs with (ReLU, None) activations and vector gate
// (models/protein_gnn.py:341-349, pretrained_model_downstream/model_kwargs.json).
using CkG0 = GvpC<64, 9, 16, 4, 9, CGVP_ACT_RELU, CGVP_ACT_NONE, 1>;
using CkG1 = GvpC<16, 4, 16, 4, 4, CGVP_ACT_RELU, CGVP_ACT_NONE, 1>;
using CkG2 = GvpC<16, 4, 16, 4, 4, CGVP_ACT_NONE, CGVP_ACT_NONE, 1>;
using ConvCk = ConvSpec<16, 4, 32, 1, CkG0, CkG1, CkG2>;

// 0: weights in shared memory, 2 CTAs/SM; 1: constant-bank weights, 2 CTAs/SM; 2: constant-bank weights, 3 CTAs/SM.
// CGVP_CONV_FWD_VARIANT overrides the default (A/B measurements).
static int conv_fwd_variant() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("CGVP_CONV_FWD_VARIANT");
        v = e ? atoi(e) : 1;
        if (v < 0 || v > 2) v = 1;
    }
    return v;
}

static bool g_fast_paths = true;
extern "C" int32_t cgvp_set_fast_paths(int32_t on) { g_fast_paths = on != 0; return 0; }
bool cgvp_fast_paths_enabled() { return g_fast_paths; }

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <class S>
static void fill_common(ConvRegArgs& a, const CgvpConvDesc* desc, const CgvpPlan* plan, const float* x_s, const float* x_v,
                        const float* e_s, const float* e_v, const float* const* h_packed) {
    memset(&a, 0, sizeof(a));
    a.E = plan->num_edges; a.N = plan->num_nodes;
    a.ntiles = (int)cdiv64(a.E, 32);
    a.mean = desc->aggr == CGVP_AGGR_MEAN; a.edge_sorted = desc->edge_sorted;
    a.perm = plan->perm; a.src = plan->src; a.dst = plan->dst; a.rowptr = plan->rowptr;
    a.x_s = x_s; a.x_v = x_v; a.e_s = e_s; a.e_v = e_v;
    for (int k = 0; k < 3; ++k) a.wp[k] = h_packed[k];
}

// Returns 1 if this descriptor / these buffers are served by a specialised kernel (rc_out holds the result), else 0.
int64_t conv_special_stash_floats(const CgvpConvDesc* desc) { return (g_fast_paths && ConvCk::matches(*desc)) ? ConvCk::STASH : 0; }

int64_t conv_special_node_floats(const CgvpConvDesc* desc) { return ConvCk::matches(*desc) ? 6 * ConvCk::G0::SO : 0; }

int conv_fwd_special(const CgvpConvDesc* desc, const CgvpPlan* plan, const float* x_s, const float* x_v, const float* e_s,
                     const float* e_v, const float* const* h_packed, float* out_s, float* out_v, float* part_head,
                     float* part_tail, float* node_ws, float* stash, cudaStream_t st, int* rc_out) {
    using S = ConvCk;
    if (!g_fast_paths || !S::matches(*desc) || plan->num_edges <= 0 || plan->num_nodes <= 0) return 0;
    if (!(aligned16(x_s) && aligned16(x_v) && aligned16(e_s) && aligned16(out_s) && aligned16(out_v) && node_ws && aligned16(node_ws)))
        return 0;
    ConvRegArgs a;
    fill_common<S>(a, desc, plan, x_s, x_v, e_s, e_v, h_packed);
    a.out_s = out_s; a.out_v = out_v; a.part_head = part_head; a.part_tail = part_tail;
    a.stash = (stash && aligned16(stash)) ? stash : nullptr;
    a.P = node_ws;
    *rc_out = 0;
    conv_node_proj_kernel<S><<<(unsigned)cdiv64(a.N, 128), 128, 0, st>>>(a);
    const int sms = cgvp_num_sms();
    const int variant = conv_fwd_variant();
    const int per_sm = variant == 2 ? 3 : 2;
    const int grid = (int)min((long long)cdiv(a.ntiles, CR_WARPS), (long long)sms * per_sm);
    const size_t smem = S::smem_fwd(variant != 0);
    cudaError_t e = cudaSuccess;
    if (variant != 0) {
        using G0 = typename S::G0; using G1 = typename S::G1; using G2 = typename S::G2;
        const int off[3] = {S::WF0, S::WF1, S::WF2}, cnt[3] = {G0::FWD_FLOATS, G1::FWD_FLOATS, G2::FWD_FLOATS};
        for (int k = 0; k < 3 && e == cudaSuccess; ++k)
            e = cudaMemcpyToSymbolAsync(c_conv_fw, a.wp[k], (size_t)cnt[k] * 4, (size_t)off[k] * 4, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) { cgvp_set_error("conv_fwd_reg: constant weights: %s", cudaGetErrorString(e)); *rc_out = (int)e; return 1; }
    }
    cgvp_prof_begin(CGVP_K_CONV_FWD, st);
    if (variant == 0) {
        e = cudaFuncSetAttribute(conv_fwd_reg_kernel<S, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) conv_fwd_reg_kernel<S, 0, 2><<<grid, CR_THREADS, smem, st>>>(a);
    } else if (variant == 1) {
        conv_fwd_reg_kernel<S, 1, 2><<<grid, CR_THREADS, smem, st>>>(a);
    } else {
        conv_fwd_reg_kernel<S, 1, 3><<<grid, CR_THREADS, smem, st>>>(a);
    }
    cgvp_prof_end(CGVP_K_CONV_FWD, st);
    if (e != cudaSuccess) { cgvp_set_error("conv_fwd_reg: %s", cudaGetErrorString(e)); *rc_out = (int)e; return 1; }
    const long long tot = a.N * S::CH;
    conv_fixup_kernel<<<(unsigned)cdiv64(tot, 256), 256, 0, st>>>(a.N, S::CH, S::SO, a.rowptr, a.mean, 5, part_head, part_tail, out_s, out_v);
    e = cudaGetLastError();
    if (e != cudaSuccess) { cgvp_set_error("launch of conv_fwd_reg_kernel failed: %s", cudaGetErrorString(e)); *rc_out = (int)e; }
    return 1;
}

int conv_bwd_special_partial_floats(const CgvpConvDesc* desc) { return ConvCk::matches(*desc) ? ConvCk::PF : 0; }

// d_x_* receive the TARGET-side gradient here; the caller adds the source side (segment reduce of dj over the source
// CSR view) and reduces the `grid_out` partials of `pf_out` floats each.
int conv_bwd_special(const CgvpConvDesc* desc, const CgvpPlan* plan, const float* x_s, const float* x_v, const float* e_s,
                     const float* e_v, const float* const* h_packed, const float* d_out_s, const float* d_out_v,
                     float* d_x_s, float* d_x_v, float* d_e_s, float* d_e_v, int accumulate_edge, float* part_head,
                     float* part_tail, float* dj, float* partial, int max_grid, float* node_ws, const float* stash, cudaStream_t st,
                     int* grid_out, int* rc_out) {
    using S = ConvCk;
    if (!g_fast_paths || !S::matches(*desc) || plan->num_edges <= 0 || plan->num_nodes <= 0) return 0;
    if (!(aligned16(x_s) && aligned16(x_v) && aligned16(e_s) && aligned16(d_out_s) && aligned16(d_out_v) && aligned16(d_x_s) &&
          aligned16(d_x_v) && aligned16(dj) && (!d_e_s || aligned16(d_e_s)) && node_ws && aligned16(node_ws) && plan->sperm &&
          plan->srowptr && max_grid >= 2))
        return 0;
    ConvRegArgs a;
    fill_common<S>(a, desc, plan, x_s, x_v, e_s, e_v, h_packed);
    // node-level workspace: [Ri | Rj | P], N x SO / SO / 2 SO floats
    a.Ri = node_ws; a.Rj = node_ws + a.N * S::G0::SO; a.P = node_ws + 2 * a.N * S::G0::SO;
    a.out_s = a.Ri; a.out_v = d_x_v; a.part_head = part_head; a.part_tail = part_tail;      // target side: [ds' ; dV_i]
    a.d_out_s = d_out_s; a.d_out_v = d_out_v; a.d_e_s = d_e_s; a.d_e_v = d_e_v; a.dj = dj; a.acc_edge = accumulate_edge;
    a.partial = partial;
    a.stash = (stash && aligned16(stash)) ? const_cast<float*>(stash) : nullptr;
    *rc_out = 0;
    const int sms = cgvp_num_sms();
    const int node_grid_full = (int)cdiv64(a.N, 128);
    int grid = (int)min((long long)cdiv(a.ntiles, CR_WARPS), (long long)sms);
    if (grid > max_grid - 1) grid = max_grid - 1;
    if (!a.stash) conv_node_proj_kernel<S><<<(unsigned)node_grid_full, 128, 0, st>>>(a);     // the recompute path needs P
    static int bvar = -1;
    if (bvar < 0) { const char* ev = getenv("CGVP_CONV_BWD_VARIANT"); bvar = ev ? (atoi(ev) != 0) : 1; }
    const size_t smem = S::smem_bwd(bvar != 0);
    cudaError_t e = bvar ? cudaFuncSetAttribute(conv_bwd_reg_kernel<S, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                         : cudaFuncSetAttribute(conv_bwd_reg_kernel<S, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (bvar) {
        using G0 = typename S::G0; using G1 = typename S::G1; using G2 = typename S::G2;
        const int off[3] = {S::WT0, S::WT1, S::WT2}, cnt[3] = {G0::TOTAL_FLOATS, G1::TOTAL_FLOATS, G2::TOTAL_FLOATS};
        for (int k = 0; k < 3 && e == cudaSuccess; ++k)
            e = cudaMemcpyToSymbolAsync(c_conv_bw, a.wp[k], (size_t)cnt[k] * 4, (size_t)off[k] * 4, cudaMemcpyDeviceToDevice, st);
    }
    if (e != cudaSuccess) { cgvp_set_error("conv_bwd_reg: %s", cudaGetErrorString(e)); *rc_out = (int)e; return 1; }
    cgvp_prof_begin(CGVP_K_CONV_BWD, st);
    if (bvar) conv_bwd_reg_kernel<S, 1><<<grid, CR_THREADS, smem, st>>>(a);
    else conv_bwd_reg_kernel<S, 0><<<grid, CR_THREADS, smem, st>>>(a);
    cgvp_prof_end(CGVP_K_CONV_BWD, st);
    const long long tot = a.N * S::CHX;
    conv_fixup_kernel<<<(unsigned)cdiv64(tot, 256), 256, 0, st>>>(a.N, S::CHX, S::NS, a.rowptr, 0, 5, part_head, part_tail, a.Ri, d_x_v);
    e = cudaGetLastError();
    if (e != cudaSuccess) { cgvp_set_error("launch of conv_bwd_reg_kernel failed: %s", cudaGetErrorString(e)); *rc_out = (int)e; return 1; }
    // source side: Rj = sum over out-edges of ds', d_x_v += sum over out-edges of dV_j (deterministic, source CSR view)
    e = cudaMemsetAsync(a.Rj, 0, (size_t)a.N * S::G0::SO * 4, st);
    if (e != cudaSuccess) { cgvp_set_error("conv_bwd_reg: %s", cudaGetErrorString(e)); *rc_out = (int)e; return 1; }
    int rc = cgvp_segment_reduce_split(dj, S::CHX, plan->srowptr, plan->sperm, a.N, CGVP_AGGR_SUM, 1, a.Rj, S::NS, d_x_v, 3 * S::NV, st);
    if (rc) { *rc_out = rc; return 1; }
    // node level: d_x_s and the node-scalar / bias rows of dW_s (one extra row of `partial` per CTA, capped by the caller's room)
    int node_grid = node_grid_full < max_grid - grid ? node_grid_full : max_grid - grid;
    if (node_grid > 2 * sms) node_grid = 2 * sms;
    a.out_s = d_x_s;
    a.part_row0 = grid;
    conv_node_post_kernel<S><<<(unsigned)node_grid, 128, 0, st>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) { cgvp_set_error("launch of conv_node_post_kernel failed: %s", cudaGetErrorString(e)); *rc_out = (int)e; return 1; }
    *grid_out = grid + node_grid;
    return 1;
}

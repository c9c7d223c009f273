// Fused GVPConv: gather -> stacked message GVPs -> deterministic per-target aggregation, and its backward.
// Replaces GVPConv.forward/message + PyG propagate (models/gvp_layers.py:291-308); nothing of size E is written in
// the forward pass.  Edges are processed in dst-sorted order in tiles of R edges; one thread owns one edge.
#include "cgvp_tile.cuh"

using namespace cgvp;

int conv_fwd_special(const CgvpConvDesc* desc, const CgvpPlan* plan, const float* x_s, const float* x_v, const float* e_s,
                     const float* e_v, const float* const* h_packed, float* out_s, float* out_v, float* part_head,
                     float* part_tail, float* node_ws, float* stash, cudaStream_t st, int* rc_out);
int64_t conv_special_node_floats(const CgvpConvDesc* desc);
int64_t conv_special_stash_floats(const CgvpConvDesc* desc);
int conv_bwd_special(const CgvpConvDesc* desc, const CgvpPlan* plan, const float* x_s, const float* x_v, const float* e_s,
                     const float* e_v, const float* const* h_packed, const float* d_out_s, const float* d_out_v,
                     float* d_x_s, float* d_x_v, float* d_e_s, float* d_e_v, int accumulate_edge, float* part_head,
                     float* part_tail, float* dj, float* partial, int max_grid, float* node_ws, const float* stash, cudaStream_t st,
                     int* grid_out, int* rc_out);

int64_t conv_tc_workspace_bytes(const CgvpConvDesc* desc, int64_t E, int64_t N);
int conv_fwd_tc(const CgvpConvDesc* desc, const CgvpPlan* plan, const float* x_s, const float* x_v, const float* e_s,
                const float* e_v, const float* const* h_packed, float* out_s, float* out_v, void* tcws, int64_t tcws_bytes,
                cudaStream_t st, int* rc_out);

int cgvp_segment_reduce_split(const float* rows, int width, const int* rowptr, const int* index, int64_t N, int aggr,
                              int beta, float* out_a, int wa, float* out_b, int wb, cudaStream_t st);

struct ConvK {
    int ns, nv, es, ev, n_gvp, aggr, edge_sorted;
    int so, vo;                           // message dims = output node dims
    GvpP g[CGVP_MAX_CHAIN];
    ChainCols cc;
    // backward
    int gs[CGVP_MAX_CHAIN + 1], gv[CGVP_MAX_CHAIN + 1], dg[CGVP_MAX_CHAIN], dvh[CGVP_MAX_CHAIN];
    DwPlan dw;
    int goff[CGVP_MAX_CHAIN], partial_floats;
    // geometry
    int R, rp, ncols, w_smem, woff[CGVP_MAX_CHAIN], wtotal;
    int team, threads;                    // threads per row (power of two) and per CTA (R * team <= threads)
    const float* wp[CGVP_MAX_CHAIN];
    long long E, N;
    int ntiles;
    // plan + tensors
    const int *perm, *src, *dst, *rowptr;
    const float *x_s, *x_v, *e_s, *e_v;
    float *out_s, *out_v;                 // forward output / backward: target-side node gradient
    float *part_head, *part_tail;         // [ntiles][so + 3 vo] partial sums of segments that straddle tiles
    int* counters;                        // [ntiles] arrival counters (zero on entry, zero on exit)
    const float *d_out_s, *d_out_v;
    float *d_e_s, *d_e_v, *dj;
    int acc_edge;
    int vec_x, vec_e, vec_d;              // 16-byte alignment of x_s / e_s / d_out_s (float4 gathers allowed)
    float* partial;
};

struct TileIdx {
    int *src, *dst, *eid;
    float* scale;
};

// ---- cooperative gathers into the tile ------------------------------------------------------------------------------
__device__ __forceinline__ void gather_s(float4* T, int rp, int rv, int col4, int col0, int width,
                                         const float* __restrict__ src, const int* rowid, const float* scale,
                                         bool vec_ok) {
    if (width <= 0) return;
    if (vec_ok && ((width | col0) & 3) == 0) {
        const int w4 = width >> 2, c40 = col4 + (col0 >> 2);
        for (int i = threadIdx.x; i < rv * w4; i += blockDim.x) {
            const int r = i / w4, c = i - r * w4;
            float4 v = __ldg(reinterpret_cast<const float4*>(src + (long long)rowid[r] * width) + c);
            if (scale) { const float f = scale[r]; v.x *= f; v.y *= f; v.z *= f; v.w *= f; }
            T[(c40 + c) * rp + r] = v;
        }
    } else {
        for (int i = threadIdx.x; i < rv * width; i += blockDim.x) {
            const int r = i / width, c = i - r * width;
            float v = __ldg(src + (long long)rowid[r] * width + c);
            if (scale) v *= scale[r];
            tile_at(T, rp, r, col4, col0 + c) = v;
        }
    }
}
__device__ __forceinline__ void gather_v(float4* T, int rp, int rv, int col4, int vpc, int ch0, int nv,
                                         const float* __restrict__ src, const int* rowid, const float* scale) {
    if (nv <= 0) return;
    const int w = 3 * nv;
    for (int i = threadIdx.x; i < rv * w; i += blockDim.x) {
        const int r = i / w, j = i - r * w, ch = j / 3, p = j - 3 * ch;
        float v = __ldg(src + (long long)rowid[r] * w + j);
        if (scale) v *= scale[r];
        tile_at(T, rp, r, col4 + p * vpc, ch0 + ch) = v;
    }
}
__device__ __forceinline__ void zero_v_pad(float4* T, int rp, int rv, int col4, int vpc, int nv) {
    const int pad = vpc * 4 - nv;
    if (pad <= 0) return;
    for (int i = threadIdx.x; i < rv * 3 * pad; i += blockDim.x) {
        const int r = i / (3 * pad), j = i - r * 3 * pad, p = j / pad;
        tile_at(T, rp, r, col4 + p * vpc, nv + (j - p * pad)) = 0.f;
    }
}

__device__ __forceinline__ void load_tile_index(const ConvK& K, const TileIdx& ix, long long p0, int rv) {
    for (int r = threadIdx.x; r < rv; r += blockDim.x) {
        const long long p = p0 + r;
        const int d = K.dst[p];
        ix.src[r] = K.src[p];
        ix.dst[r] = d;
        ix.eid[r] = K.edge_sorted ? (int)p : K.perm[p];
        ix.scale[r] = K.aggr == CGVP_AGGR_MEAN ? 1.f / (float)max(K.rowptr[d + 1] - K.rowptr[d], 1) : 1.f;
    }
}

// message input (s_j, e_s, s_i), (V_j, e_V, V_i)  -- gvp_layers.py:306
__device__ __forceinline__ void stage_message_input(const ConvK& K, float4* T, const TileIdx& ix, int rv) {
    const int rp = K.rp, s0 = K.cc.s[0], v0 = K.cc.v[0], vpc = K.cc.vpc[0];
    gather_s(T, rp, rv, s0, 0, K.ns, K.x_s, ix.src, nullptr, K.vec_x);
    gather_s(T, rp, rv, s0, K.ns, K.es, K.e_s, ix.eid, nullptr, K.vec_e);
    gather_s(T, rp, rv, s0, K.ns + K.es, K.ns, K.x_s, ix.dst, nullptr, K.vec_x);
    gather_v(T, rp, rv, v0, vpc, 0, K.nv, K.x_v, ix.src, nullptr);
    gather_v(T, rp, rv, v0, vpc, K.nv, K.ev, K.e_v, ix.eid, nullptr);
    gather_v(T, rp, rv, v0, vpc, K.nv + K.ev, K.nv, K.x_v, ix.dst, nullptr);
    zero_v_pad(T, rp, rv, v0, vpc, 2 * K.nv + K.ev);
}

// ---- deterministic segmented reduction of tile rows over the (sorted) target node -----------------------------------
// Channels: ch < S -> scalar column s_col0 + ch;  else vector channel v_ch0 + (ch-S)/3, plane (ch-S)%3.
// Complete segments are written to (out_s, out_v); the (at most two) segments that straddle the tile go to the
// partial buffers, and the LAST CTA to arrive at such a node adds its pieces in tile order (no atomics on data,
// no spinning): the result does not depend on scheduling.
struct ReduceMap {
    int s_col4, s_col0, S;
    int v_col4, vpc, v_ch0, V;
};

__device__ __forceinline__ float reduce_read(const float4* T, int rp, const ReduceMap& m, int ch, int r) {
    const float* Tf = reinterpret_cast<const float*>(T);
    if (ch < m.S) {
        const int col = m.s_col0 + ch;
        return Tf[((m.s_col4 + (col >> 2)) * rp + r) * 4 + (col & 3)];
    }
    const int j = ch - m.S, c = m.v_ch0 + j / 3, p = j % 3;
    return Tf[((m.v_col4 + p * m.vpc + (c >> 2)) * rp + r) * 4 + (c & 3)];
}

__device__ __forceinline__ void segmented_reduce_tile(const ConvK& K, const float4* T, const TileIdx& ix, const ReduceMap& m,
                                                      int tile, long long p0, int rv, bool apply_scale, int* flags) {
    const int rp = K.rp, CH = m.S + 3 * m.V;
    const int n_first = ix.dst[0], n_last = ix.dst[rv - 1];
    const long long p1 = p0 + rv;
    const int span = n_last - n_first + 1;
    for (int i = threadIdx.x; i < span * CH; i += blockDim.x) {
        const int n = n_first + i / CH, ch = i % CH;
        const long long a = K.rowptr[n], b = K.rowptr[n + 1];
        const int ra = (int)(max(a, p0) - p0), rb = (int)(min(b, p1) - p0);
        if (ra >= rb) continue;
        float sum = 0.f;
        for (int r = ra; r < rb; ++r) sum += reduce_read(T, rp, m, ch, r);
        if (a >= p0 && b <= p1) {
            const float f = (apply_scale && K.aggr == CGVP_AGGR_MEAN) ? 1.f / (float)max((int)(b - a), 1) : 1.f;
            if (ch < m.S) K.out_s[(long long)n * m.S + ch] = sum * f;
            else K.out_v[(long long)n * 3 * m.V + (ch - m.S)] = sum * f;
        } else if (a < p0) {
            K.part_head[(long long)tile * CH + ch] = sum;
        } else {
            K.part_tail[(long long)tile * CH + ch] = sum;
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        flags[0] = flags[1] = -1;
        const long long a0 = K.rowptr[n_first];
        if (a0 < p0) {                                   // head node started in an earlier tile
            const int ta = (int)(a0 / K.R), tb = (int)((K.rowptr[n_first + 1] - 1) / K.R);
            if (atomicAdd(&K.counters[ta], 1) == tb - ta) flags[0] = n_first;
        }
        const long long al = K.rowptr[n_last], bl = K.rowptr[n_last + 1];
        if (al >= p0 && bl > p1) {                       // tail node continues into later tiles
            const int tb = (int)((bl - 1) / K.R);
            if (atomicAdd(&K.counters[tile], 1) == tb - tile) flags[1] = n_last;
        }
    }
    __syncthreads();
    for (int f = 0; f < 2; ++f) {
        const int n = flags[f];
        if (n < 0) continue;
        __threadfence();
        const long long a = K.rowptr[n], b = K.rowptr[n + 1];
        const int ta = (int)(a / K.R), tb = (int)((b - 1) / K.R);
        const float sc = (apply_scale && K.aggr == CGVP_AGGR_MEAN) ? 1.f / (float)max((int)(b - a), 1) : 1.f;
        for (int ch = threadIdx.x; ch < CH; ch += blockDim.x) {
            float sum = __ldcg(K.part_tail + (long long)ta * CH + ch);
            for (int t = ta + 1; t <= tb; ++t) sum += __ldcg(K.part_head + (long long)t * CH + ch);
            if (ch < m.S) K.out_s[(long long)n * m.S + ch] = sum * sc;
            else K.out_v[(long long)n * 3 * m.V + (ch - m.S)] = sum * sc;
        }
        if (threadIdx.x == 0) K.counters[ta] = 0;        // leave the counters clean for the next launch
    }
}

template <bool WS>
__device__ __forceinline__ const float* conv_weights(const ConvK& K, const float* wsm, int k) {
    if (WS) return wsm + K.woff[k];   // shared-memory resident (compile-time choice keeps these loads LDS)
    return K.wp[k];
}

struct ConvSmem {
    float* wsm;
    TileIdx ix;
    int* flags;
    float4* T;
};
__device__ __forceinline__ ConvSmem carve(const ConvK& K, unsigned char* smem) {
    ConvSmem s;
    s.wsm = reinterpret_cast<float*>(smem);
    int* ip = reinterpret_cast<int*>(s.wsm + K.wtotal);
    s.ix.src = ip; s.ix.dst = ip + K.R; s.ix.eid = ip + 2 * K.R;
    s.ix.scale = reinterpret_cast<float*>(ip + 3 * K.R);
    s.flags = ip + 4 * K.R;
    s.T = reinterpret_cast<float4*>(ip + 4 * K.R + 4);
    return s;
}

template <bool WS>
__global__ void __launch_bounds__(CGVP_MAX_THREADS) conv_fwd_kernel(const __grid_constant__ ConvK K) {
    extern __shared__ __align__(16) unsigned char smem[];
    const ConvSmem sm = carve(K, smem);
    float4* T = sm.T;
    const int rp = K.rp, L = K.n_gvp;
    if (K.w_smem)
        for (int k = 0; k < L; ++k) copy_f4(sm.wsm + K.woff[k], K.wp[k], K.g[k].fwd_floats);
    for (int i = threadIdx.x; i < K.ncols * rp; i += blockDim.x) T[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    ReduceMap m;
    m.s_col4 = K.cc.s[L]; m.s_col0 = 0; m.S = K.so;
    m.v_col4 = K.cc.v[L]; m.vpc = K.cc.vpc[L]; m.v_ch0 = 0; m.V = K.vo;
    for (int t = blockIdx.x; t < K.ntiles; t += gridDim.x) {
        const long long p0 = (long long)t * K.R;
        const int rv = (int)min((long long)K.R, K.E - p0);
        load_tile_index(K, sm.ix, p0, rv);
        __syncthreads();
        stage_message_input(K, T, sm.ix, rv);
        __syncthreads();
        {   // K.team threads per row; rows >= rv hold stale (finite) data and are never stored
            const int row = threadIdx.x / K.team, rank = threadIdx.x % K.team;
            for (int k = 0; k < L; ++k)
                gvp_fwd_team<false>(K.g[k], conv_weights<WS>(K, sm.wsm, k), T, rp, row < K.R ? row : 0, stage_io(K.cc, k), rank,
                                    K.team, row < K.R);
        }
        __syncthreads();
        segmented_reduce_tile(K, T, sm.ix, m, t, p0, rv, true, sm.flags);
        __syncthreads();
    }
}

template <int NSLOT, bool WS>
__global__ void __launch_bounds__(CGVP_MAX_THREADS) conv_bwd_kernel(const __grid_constant__ ConvK K) {
    extern __shared__ __align__(16) unsigned char smem[];
    const ConvSmem sm = carve(K, smem);
    float4* T = sm.T;
    const int rp = K.rp, L = K.n_gvp;
    if (K.w_smem)
        for (int k = 0; k < L; ++k) copy_f4(sm.wsm + K.woff[k], K.wp[k], K.g[k].total_floats);
    for (int i = threadIdx.x; i < K.ncols * rp; i += blockDim.x) T[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    float* partial = K.partial + (long long)blockIdx.x * K.partial_floats;
    DwAcc<NSLOT> dwacc;
    dwacc.init();
    ReduceMap m;                                   // target-side slice of the message-input gradient
    m.s_col4 = K.gs[0]; m.s_col0 = K.ns + K.es; m.S = K.ns;
    m.v_col4 = K.gv[0]; m.vpc = K.cc.vpc[0]; m.v_ch0 = K.nv + K.ev; m.V = K.nv;
    const int wj = K.ns + 3 * K.nv;
    for (int t = blockIdx.x; t < K.ntiles; t += gridDim.x) {
        const long long p0 = (long long)t * K.R;
        const int rv = (int)min((long long)K.R, K.E - p0);
        load_tile_index(K, sm.ix, p0, rv);
        __syncthreads();
        stage_message_input(K, T, sm.ix, rv);
        // d(message_e) = d_out[dst_e] (/ deg for mean), zero padded
        gather_s(T, rp, rv, K.gs[L], 0, K.so, K.d_out_s, sm.ix.dst, K.aggr == CGVP_AGGR_MEAN ? sm.ix.scale : nullptr, K.vec_d);
        {
            const int pad = ((K.so + 3) & ~3) - K.so;
            for (int i = threadIdx.x; i < rv * pad; i += blockDim.x) tile_at(T, rp, i / pad, K.gs[L], K.so + i % pad) = 0.f;
        }
        gather_v(T, rp, rv, K.gv[L], K.cc.vpc[L], 0, K.vo, K.d_out_v, sm.ix.dst,
                 K.aggr == CGVP_AGGR_MEAN ? sm.ix.scale : nullptr);
        zero_v_pad(T, rp, rv, K.gv[L], K.cc.vpc[L], K.vo);
        __syncthreads();
        {
            const int row = threadIdx.x / K.team, rank = threadIdx.x % K.team;
            const bool work = row < K.R;
            const int r = work ? row : 0;
            for (int k = 0; k < L; ++k)
                gvp_fwd_team<true>(K.g[k], conv_weights<WS>(K, sm.wsm, k), T, rp, r, stage_io(K.cc, k), rank, K.team, work);
            for (int k = L - 1; k >= 0; --k) {
                GradIO d;
                d.gs_in = K.gs[k + 1]; d.gv_in = K.gv[k + 1]; d.gv_in_pc = K.cc.vpc[k + 1];
                d.gs_out = K.gs[k]; d.gv_out = K.gv[k]; d.gv_out_pc = K.cc.vpc[k];
                d.dg = K.dg[k]; d.dvh = K.dvh[k]; d.dvh_pc = K.cc.vhpc[k];
                gvp_bwd_team(K.g[k], conv_weights<WS>(K, sm.wsm, k), T, rp, r, stage_io(K.cc, k), d, rank, K.team, work);
            }
        }
        __syncthreads();
        // edge-attribute gradient (one row per edge, written or accumulated)
        if (K.d_e_s) {
            for (int i = threadIdx.x; i < rv * K.es; i += blockDim.x) {
                const int r = i / K.es, c = i - r * K.es;
                float* dst = K.d_e_s + (long long)sm.ix.eid[r] * K.es + c;
                const float v = tile_at(T, rp, r, K.gs[0], K.ns + c);
                *dst = K.acc_edge ? *dst + v : v;
            }
        }
        if (K.d_e_v && K.ev > 0) {
            const int w = 3 * K.ev;
            for (int i = threadIdx.x; i < rv * w; i += blockDim.x) {
                const int r = i / w, j = i - r * w, ch = j / 3, p = j - 3 * ch;
                float* dst = K.d_e_v + (long long)sm.ix.eid[r] * w + j;
                const float v = tile_at(T, rp, r, K.gv[0] + p * K.cc.vpc[0], K.nv + ch);
                *dst = K.acc_edge ? *dst + v : v;
            }
        }
        // source-side slice, one merged row per edge (reduced over the source CSR view afterwards)
        for (int i = threadIdx.x; i < rv * wj; i += blockDim.x) {
            const int r = i / wj, c = i - r * wj;
            float v;
            if (c < K.ns) v = tile_at(T, rp, r, K.gs[0], c);
            else { const int j = c - K.ns, ch = j / 3, p = j - 3 * ch; v = tile_at(T, rp, r, K.gv[0] + p * K.cc.vpc[0], ch); }
            K.dj[(p0 + r) * wj + c] = v;
        }
        segmented_reduce_tile(K, T, sm.ix, m, t, p0, rv, false, sm.flags);
        dwacc.tile(K.dw, T, rp, rv, partial);
        __syncthreads();
    }
    dwacc.flush(K.dw, partial);
}

// ---- host side --------------------------------------------------------------------------------------------------------
static int build_conv_k(const CgvpConvDesc* desc, bool backward, ConvK& K) {
    memset(&K, 0, sizeof(K));
    CGVP_REQUIRE(desc, "conv: null descriptor");
    CGVP_REQUIRE(desc->n_gvp >= 1 && desc->n_gvp <= CGVP_MAX_CHAIN, "conv: n_gvp %d out of range", desc->n_gvp);
    CGVP_REQUIRE(desc->ns > 0 && desc->nv >= 0 && desc->es >= 0 && desc->ev >= 0, "conv: bad dims");
    CGVP_REQUIRE(desc->aggr == CGVP_AGGR_SUM || desc->aggr == CGVP_AGGR_MEAN, "conv: bad aggr");
    K.ns = desc->ns; K.nv = desc->nv; K.es = desc->es; K.ev = desc->ev; K.n_gvp = desc->n_gvp; K.aggr = desc->aggr;
    K.edge_sorted = desc->edge_sorted;
    int cs = 2 * K.ns + K.es, cv = 2 * K.nv + K.ev;
    for (int k = 0; k < K.n_gvp; ++k) {
        if (cgvp_validate_gvp(desc->gvp[k], "conv")) return -1;
        K.g[k] = make_gvp_p(desc->gvp[k]);
        CGVP_REQUIRE(K.g[k].si == cs && K.g[k].vi == cv, "conv: message GVP %d expects (%d,%d) but receives (%d,%d)", k,
                     K.g[k].si, K.g[k].vi, cs, cv);
        cs = K.g[k].so; cv = K.g[k].vo;
    }
    K.so = cs; K.vo = cv;
    const int L = K.n_gvp;
    int col = 0;
    K.cc = plan_chain_cols(K.g, L, backward, backward, col);
    col += K.cc.ncols;
    if (backward) {
        for (int k = 0; k <= L; ++k) {
            K.gs[k] = col; col += k < L ? K.g[k].ksd4 : K.g[L - 1].so4;
            K.gv[k] = col; col += 3 * K.cc.vpc[k];
        }
        for (int k = 0; k < L; ++k) { K.dg[k] = col; col += K.g[k].vo4; K.dvh[k] = col; col += 3 * K.g[k].h4; }
        int goff = 0;
        for (int k = 0; k < L; ++k) {
            K.goff[k] = goff;
            GradIO d;
            d.gs_in = K.gs[k + 1]; d.gv_in = K.gv[k + 1]; d.gv_in_pc = K.cc.vpc[k + 1];
            d.gs_out = K.gs[k]; d.gv_out = K.gv[k]; d.gv_out_pc = K.cc.vpc[k];
            d.dg = K.dg[k]; d.dvh = K.dvh[k]; d.dvh_pc = K.cc.vhpc[k];
            dw_add_gvp(K.dw, K.g[k], stage_io(K.cc, k), d, goff);
            goff += K.g[k].fwd_floats;
        }
        K.partial_floats = (int)align_up(goff, 4);
    }
    K.ncols = col;
    int wt = 0;
    for (int k = 0; k < L; ++k) {
        K.woff[k] = wt;
        wt += (int)align_up(backward ? K.g[k].total_floats : K.g[k].fwd_floats, 4);
    }
    K.wtotal = wt;
    return 0;
}

// rows per tile: the largest multiple of 32 (<= 128) whose tile fits; prefer two resident CTAs per SM in forward
static int conv_geometry(ConvK& K, int smem_max, bool backward, size_t* smem_bytes, int* ctas_per_sm) {
    const int wt_full = K.wtotal;
    for (int pass = 0; pass < 2; ++pass) {
        const int wbytes = pass == 0 ? wt_full * 4 : 0;
        const auto need = [&](int R) { return (size_t)wbytes + (size_t)(4 * R + 4) * 4 + (size_t)K.ncols * (R + 1) * 16 + 16; };
        int best_r = 0, best_c = 0;
        for (int R = CGVP_THREADS; R >= 8; R = R > 32 ? R - 32 : R / 2) {   // 128, 96, 64, 32, 16, 8
            const size_t n = need(R);
            if (n > (size_t)smem_max) continue;
            int c = (int)((size_t)(228 * 1024) / (n + 1024));
            if (c < 1) c = 1;
            if (c > 4) c = 4;
            // score = concurrently resident rows per SM
            if (R * c > best_r * best_c) { best_r = R; best_c = c; }
        }
        if (best_r > 0) {
            K.R = best_r; K.rp = best_r + 1; K.w_smem = pass == 0;
            // few rows fit at wide dims: up to 8 threads team up on each row (more warps to hide latency, see gvp_fwd_team)
            K.team = 1;
            if ((best_r & (best_r - 1)) == 0 && best_r <= 64)
                while (K.team < 8 && best_r * K.team * 2 <= CGVP_MAX_THREADS) K.team *= 2;
            K.threads = best_r * K.team < CGVP_THREADS ? CGVP_THREADS : best_r * K.team;
            if (K.team == 1) K.threads = CGVP_THREADS;
            K.wtotal = pass == 0 ? wt_full : 0;
            *smem_bytes = need(best_r);
            *ctas_per_sm = best_c;
            return 0;
        }
    }
    cgvp_set_error("conv: tile of %d float4 columns does not fit in shared memory", K.ncols);
    return -1;
}

// `node_floats`: per-node scratch of the specialised kernels (per-node projections / reductions of message GVP 0), in floats
// per node; it sits at the end so that the other offsets do not depend on it.
static int64_t conv_ws_layout(const ConvK& K, int64_t E, int64_t N, bool backward, int grid, int64_t* o_head,
                              int64_t* o_tail, int64_t* o_cnt, int64_t* o_dj, int64_t* o_partial, int64_t* o_reduced,
                              int64_t node_floats = 0, int64_t* o_node = nullptr) {
    const int64_t ntiles_max = cdiv64(E > 0 ? E : 1, 8);
    const int CH = (K.so + 3 * K.vo) > (K.ns + 3 * K.nv) ? (K.so + 3 * K.vo) : (K.ns + 3 * K.nv);
    int64_t off = 0;
    *o_head = off; off += align_up(ntiles_max * CH * 4, 256);
    *o_tail = off; off += align_up(ntiles_max * CH * 4, 256);
    *o_cnt = off; off += align_up(ntiles_max * 4, 256);
    *o_dj = off;
    if (backward) off += align_up(E * (K.ns + 3 * K.nv) * 4, 256);
    *o_partial = off;
    if (backward) off += align_up((int64_t)grid * K.partial_floats * 4, 256);
    *o_reduced = off;
    if (backward) off += align_up((int64_t)K.partial_floats * 4, 256);
    if (o_node) *o_node = off;
    off += align_up((N > 0 ? N : 1) * node_floats * 4, 256);
    return off + 256;
}

extern "C" int64_t cgvp_conv_workspace_bytes(const CgvpConvDesc* desc, int64_t num_edges, int64_t num_nodes,
                                             int32_t backward) {
    ConvK K;
    if (build_conv_k(desc, backward != 0, K)) return -1;
    const int sms = cgvp_num_sms() > 0 ? cgvp_num_sms() : 148;
    int64_t a, b, c, d, e, f;
    // the tensor-core forward (conv_tc.cu) keeps its own region after the common layout
    return conv_ws_layout(K, num_edges, num_nodes, backward != 0, 4 * sms, &a, &b, &c, &d, &e, &f, conv_special_node_floats(desc)) +
           (backward ? 0 : conv_tc_workspace_bytes(desc, num_edges, num_nodes));
}

static int conv_common_checks(const ConvK& K, const CgvpPlan* plan, const float* x_s, const float* x_v, const float* e_s,
                              const float* e_v, const float* const* h_packed) {
    CGVP_REQUIRE(plan && plan->num_edges >= 0 && plan->num_nodes >= 0, "conv: null plan");
    if (plan->num_edges > 0 && plan->num_nodes > 0)
        CGVP_REQUIRE(x_s && (K.nv == 0 || x_v) && (K.es == 0 || e_s) && (K.ev == 0 || e_v), "conv: null input tensor");
    CGVP_REQUIRE(h_packed, "conv: packed weights missing");
    for (int k = 0; k < K.n_gvp; ++k)
        CGVP_REQUIRE(h_packed[k] && ((uintptr_t)h_packed[k] & 15) == 0, "conv: packed block %d null/unaligned", k);
    if (plan->num_edges > 0)
        CGVP_REQUIRE(plan->perm && plan->src && plan->dst && plan->rowptr, "conv: plan buffers missing");
    return 0;
}

extern "C" int64_t cgvp_conv_stash_bytes(const CgvpConvDesc* desc, int64_t num_edges) {
    if (!desc || num_edges <= 0) return 0;
    return conv_special_stash_floats(desc) * num_edges * 4;
}

extern "C" int32_t cgvp_conv_fwd(const CgvpConvDesc* desc, const CgvpPlan* plan, const float* x_s, const float* x_v,
                                 const float* e_s, const float* e_v, const float* const* h_packed, float* out_s,
                                 float* out_v, void* ws, int64_t ws_bytes, cgvp_stream_t stream) {
    return cgvp_conv_fwd_stash(desc, plan, x_s, x_v, e_s, e_v, h_packed, out_s, out_v, ws, ws_bytes, nullptr, stream);
}

extern "C" int32_t cgvp_conv_fwd_stash(const CgvpConvDesc* desc, const CgvpPlan* plan, const float* x_s, const float* x_v,
                                       const float* e_s, const float* e_v, const float* const* h_packed, float* out_s,
                                       float* out_v, void* ws, int64_t ws_bytes, void* stash, cgvp_stream_t stream) {
    ConvK K;
    if (build_conv_k(desc, false, K)) return -1;
    if (conv_common_checks(K, plan, x_s, x_v, e_s, e_v, h_packed)) return -1;
    CGVP_REQUIRE(out_s && (K.vo == 0 || out_v), "conv_fwd: null output");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t E = plan->num_edges, N = plan->num_nodes;
    if (E > 0 && N > 0) {   // specialised kernels for the dims they were compiled for (conv_tc.cu, conv_reg.cu)
        int64_t oh, ot, oc, od, op, orr, on;
        const int64_t need = conv_ws_layout(K, E, N, false, 4 * (cgvp_num_sms() > 0 ? cgvp_num_sms() : 148), &oh, &ot, &oc, &od, &op, &orr,
                                            conv_special_node_floats(desc), &on);
        if (ws && ws_bytes >= need && ((uintptr_t)ws & 255) == 0) {
            char* base = reinterpret_cast<char*>(ws);
            int rc = 0;
            if (conv_fwd_tc(desc, plan, x_s, x_v, e_s, e_v, h_packed, out_s, out_v, base + need, ws_bytes - need, st, &rc)) return rc;
            if (conv_fwd_special(desc, plan, x_s, x_v, e_s, e_v, h_packed, out_s, out_v, reinterpret_cast<float*>(base + oh),
                                 reinterpret_cast<float*>(base + ot), reinterpret_cast<float*>(base + on),
                                 reinterpret_cast<float*>(stash), st, &rc))
                return rc;
        }
    }
    // nodes without incoming edges receive 0 (PyG scatter with dim_size = N)
    if (N > 0) {
        CGVP_CUDA(cudaMemsetAsync(out_s, 0, (size_t)N * K.so * 4, st));
        if (K.vo > 0) CGVP_CUDA(cudaMemsetAsync(out_v, 0, (size_t)N * K.vo * 12, st));
    }
    if (E == 0 || N == 0) return 0;
    size_t smem = 0;
    int per_sm = 1;
    if (conv_geometry(K, cgvp_max_smem_optin(), false, &smem, &per_sm)) return -1;
    K.E = E; K.N = N;
    K.ntiles = (int)cdiv64(E, K.R);
    const int sms = cgvp_num_sms();
    int grid = K.ntiles < sms * per_sm ? K.ntiles : sms * per_sm;
    int64_t oh, ot, oc, od, op, orr;
    const int64_t need = conv_ws_layout(K, E, N, false, grid, &oh, &ot, &oc, &od, &op, &orr);
    CGVP_REQUIRE(ws && ws_bytes >= need && ((uintptr_t)ws & 255) == 0, "conv_fwd: workspace too small or unaligned (%lld < %lld)",
                 (long long)ws_bytes, (long long)need);
    char* base = reinterpret_cast<char*>(ws);
    K.part_head = reinterpret_cast<float*>(base + oh);
    K.part_tail = reinterpret_cast<float*>(base + ot);
    K.counters = reinterpret_cast<int*>(base + oc);
    CGVP_CUDA(cudaMemsetAsync(K.counters, 0, (size_t)K.ntiles * 4, st));
    K.perm = plan->perm; K.src = plan->src; K.dst = plan->dst; K.rowptr = plan->rowptr;
    K.x_s = x_s; K.x_v = x_v; K.e_s = e_s; K.e_v = e_v; K.out_s = out_s; K.out_v = out_v;
    K.vec_x = ((uintptr_t)x_s & 15) == 0; K.vec_e = ((uintptr_t)e_s & 15) == 0;
    for (int k = 0; k < K.n_gvp; ++k) K.wp[k] = h_packed[k];
    cgvp_prof_begin(CGVP_K_CONV_FWD, st);
    if (K.w_smem) {
        CGVP_CUDA(cudaFuncSetAttribute(conv_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv_fwd_kernel<true><<<grid, K.threads, smem, st>>>(K);
    } else {
        CGVP_CUDA(cudaFuncSetAttribute(conv_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conv_fwd_kernel<false><<<grid, K.threads, smem, st>>>(K);
    }
    cgvp_prof_end(CGVP_K_CONV_FWD, st);
    CGVP_LAUNCH_CHECK("conv_fwd_kernel");
    return 0;
}

extern "C" int32_t cgvp_conv_bwd(const CgvpConvDesc* desc, const CgvpPlan* plan, const float* x_s, const float* x_v,
                                 const float* e_s, const float* e_v, const float* const* h_packed, const float* d_out_s,
                                 const float* d_out_v, float* d_x_s, float* d_x_v, float* d_e_s, float* d_e_v,
                                 int32_t accumulate_edge, float* const* h_packed_grads, void* ws, int64_t ws_bytes,
                                 cgvp_stream_t stream) {
    return cgvp_conv_bwd_stash(desc, plan, x_s, x_v, e_s, e_v, h_packed, d_out_s, d_out_v, d_x_s, d_x_v, d_e_s, d_e_v,
                               accumulate_edge, h_packed_grads, ws, ws_bytes, nullptr, stream);
}

extern "C" int32_t cgvp_conv_bwd_stash(const CgvpConvDesc* desc, const CgvpPlan* plan, const float* x_s, const float* x_v,
                                       const float* e_s, const float* e_v, const float* const* h_packed, const float* d_out_s,
                                       const float* d_out_v, float* d_x_s, float* d_x_v, float* d_e_s, float* d_e_v,
                                       int32_t accumulate_edge, float* const* h_packed_grads, void* ws, int64_t ws_bytes,
                                       const void* stash, cgvp_stream_t stream) {
    ConvK K;
    if (build_conv_k(desc, true, K)) return -1;
    if (conv_common_checks(K, plan, x_s, x_v, e_s, e_v, h_packed)) return -1;
    CGVP_REQUIRE(d_out_s && (K.vo == 0 || d_out_v), "conv_bwd: null upstream gradient");
    CGVP_REQUIRE(d_x_s && (K.nv == 0 || d_x_v), "conv_bwd: null node gradient output");
    CGVP_REQUIRE(h_packed_grads, "conv_bwd: packed gradient blocks missing");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t E = plan->num_edges, N = plan->num_nodes;
    if (E > 0 && N > 0 && plan->sperm && plan->srowptr) {   // specialised register-resident kernel (conv_reg.cu)
        const int sms_ = cgvp_num_sms();
        int64_t oh, ot, oc, od, op, orr, on;
        // partial rows: one per CTA of the edge kernel (<= SMs) + one per CTA of the node-level kernel (<= 2 SMs)
        const int64_t need = conv_ws_layout(K, E, N, true, 3 * sms_, &oh, &ot, &oc, &od, &op, &orr, conv_special_node_floats(desc), &on);
        if (ws && ws_bytes >= need && ((uintptr_t)ws & 255) == 0) {
            char* base = reinterpret_cast<char*>(ws);
            float* dj = reinterpret_cast<float*>(base + od);
            float* partial = reinterpret_cast<float*>(base + op);
            int rc = 0, grid = 0;
            // the specialised path does everything up to the per-CTA partials: edge kernel, target / source side reductions,
            // node-level finish
            if (conv_bwd_special(desc, plan, x_s, x_v, e_s, e_v, h_packed, d_out_s, d_out_v, d_x_s, d_x_v, d_e_s, d_e_v,
                                 accumulate_edge, reinterpret_cast<float*>(base + oh), reinterpret_cast<float*>(base + ot), dj,
                                 partial, 3 * sms_, reinterpret_cast<float*>(base + on), reinterpret_cast<const float*>(stash), st,
                                 &grid, &rc)) {
                if (rc) return rc;
                CgvpSeg seg[CGVP_MAX_SEGS];
                memset(seg, 0, sizeof(seg));
                for (int k = 0; k < K.n_gvp; ++k) { seg[k].dst = h_packed_grads[k]; seg[k].off = K.goff[k]; seg[k].n = K.g[k].fwd_floats; }
                return cgvp_reduce_partials(partial, grid, K.partial_floats, reinterpret_cast<float*>(base + orr), seg, K.n_gvp, st);
            }
        }
    }
    if (N > 0) {
        CGVP_CUDA(cudaMemsetAsync(d_x_s, 0, (size_t)N * K.ns * 4, st));
        if (K.nv > 0) CGVP_CUDA(cudaMemsetAsync(d_x_v, 0, (size_t)N * K.nv * 12, st));
    }
    size_t smem = 0;
    int per_sm = 1;
    if (conv_geometry(K, cgvp_max_smem_optin(), true, &smem, &per_sm)) return -1;
    K.E = E; K.N = N;
    K.ntiles = (int)cdiv64(E, K.R);
    const int sms = cgvp_num_sms();
    int grid = K.ntiles < sms * per_sm ? K.ntiles : sms * per_sm;
    if (grid < 1) grid = 1;
    int64_t oh, ot, oc, od, op, orr;
    const int64_t need = conv_ws_layout(K, E, N, true, grid, &oh, &ot, &oc, &od, &op, &orr);
    CGVP_REQUIRE(ws && ws_bytes >= need && ((uintptr_t)ws & 255) == 0, "conv_bwd: workspace too small or unaligned (%lld < %lld)",
                 (long long)ws_bytes, (long long)need);
    char* base = reinterpret_cast<char*>(ws);
    K.part_head = reinterpret_cast<float*>(base + oh);
    K.part_tail = reinterpret_cast<float*>(base + ot);
    K.counters = reinterpret_cast<int*>(base + oc);
    K.dj = reinterpret_cast<float*>(base + od);
    K.partial = reinterpret_cast<float*>(base + op);
    float* reduced = reinterpret_cast<float*>(base + orr);
    CGVP_CUDA(cudaMemsetAsync(K.partial, 0, (size_t)grid * K.partial_floats * 4, st));
    if (E > 0 && N > 0) {
        CGVP_REQUIRE(plan->sperm && plan->srowptr, "conv_bwd: source CSR view missing from the plan");
        CGVP_CUDA(cudaMemsetAsync(K.counters, 0, (size_t)K.ntiles * 4, st));
        K.perm = plan->perm; K.src = plan->src; K.dst = plan->dst; K.rowptr = plan->rowptr;
        K.x_s = x_s; K.x_v = x_v; K.e_s = e_s; K.e_v = e_v;
        K.vec_x = ((uintptr_t)x_s & 15) == 0; K.vec_e = ((uintptr_t)e_s & 15) == 0; K.vec_d = ((uintptr_t)d_out_s & 15) == 0;
        K.out_s = d_x_s; K.out_v = d_x_v;
        K.d_out_s = d_out_s; K.d_out_v = d_out_v; K.d_e_s = d_e_s; K.d_e_v = d_e_v; K.acc_edge = accumulate_edge;
        for (int k = 0; k < K.n_gvp; ++k) K.wp[k] = h_packed[k];
        cgvp_prof_begin(CGVP_K_CONV_BWD, st);
        if (K.w_smem) {
            CGVP_CUDA(cudaFuncSetAttribute(conv_bwd_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            conv_bwd_kernel<2, true><<<grid, K.threads, smem, st>>>(K);
        } else {
            CGVP_CUDA(cudaFuncSetAttribute(conv_bwd_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            conv_bwd_kernel<2, false><<<grid, K.threads, smem, st>>>(K);
        }
        cgvp_prof_end(CGVP_K_CONV_BWD, st);
        CGVP_LAUNCH_CHECK("conv_bwd_kernel");
        // d_x += sum over outgoing edges of the source-side slice (deterministic, source CSR view)
        const int rc = cgvp_segment_reduce_split(K.dj, K.ns + 3 * K.nv, plan->srowptr, plan->sperm, N, CGVP_AGGR_SUM, 1,
                                                 d_x_s, K.ns, d_x_v, 3 * K.nv, st);
        if (rc) return rc;
    }
    CgvpSeg seg[CGVP_MAX_SEGS];
    memset(seg, 0, sizeof(seg));
    for (int k = 0; k < K.n_gvp; ++k) { seg[k].dst = h_packed_grads[k]; seg[k].off = K.goff[k]; seg[k].n = K.g[k].fwd_floats; }
    return cgvp_reduce_partials(K.partial, grid, K.partial_floats, reduced, seg, K.n_gvp, st);
}

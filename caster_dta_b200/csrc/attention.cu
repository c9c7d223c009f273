// Multi-head cross-attention core on PACKED rows (no padding): the scaled-dot-product stage of the two
// nn.MultiheadAttention modules of CrossAttentionModule (models/joint_gnn.py:350-361), residues <-> atoms.
//
//   q:[Nq, H*D]  k, v:[Nk, H*D]   graph g owns query rows [qptr[g], qptr[g+1]) and key rows [kptr[g], kptr[g+1])
//   out[i, h] = sum_j softmax_j(scale * q[i,h] . k[j,h]) v[j,h]      over the keys j of the query's own graph
//   weights[g, i - qptr[g], j - kptr[g]] = mean_h softmax(...)        (what need_weights=True returns, :355)
//
// The reference pads both sides to [B, max, E], materialises [B, H, Lq, Lk] scores, masks, soft-maxes, averages and
// multiplies -- ~10 HBM round trips over the score tensor per direction and as many again in the backward.  Here one warp
// (or the 8 warps of a CTA when rows are few and streams long) owns one query row (forward, dQ) or one key row (dK/dV) and
// streams the other side's rows of the same graph with fully coalesced loads; soft-max statistics (row max, row sum) are
// kept per (row, head) for the backward, which recomputes the probabilities instead of storing them.  fp32 throughout,
// deterministic (fixed summation order, no atomics).
#include "cgvp_common.cuh"

struct AttnArgs {
    const float *q, *k, *v;
    const int64_t *qptr, *kptr, *qbatch, *kbatch;
    long long Nq, Nk, B;
    float scale;
    int lq_max, lk_max;
    // forward
    const float* q_fill;        // [H*D] query of every PADDED row of the reference (or NULL)
    float *out, *stats, *weights, *w_fill;
    // backward
    const float *o, *d_out;
    float *dsum, *dq, *dk, *dv;
};

// Lane map: a warp covers one packed row of E = H * D floats with ONE coalesced load per lane -- lane owns VEC = E / 32
// consecutive floats, LPH = 32 / H lanes share a head.  Dot products are VEC-wide partial sums finished with log2(LPH)
// shuffles; the probability-weighted sums need no reduction (each lane accumulates its own VEC output columns).
// WPR warps share one row (WPR = 8: the whole CTA; used for few rows with long streams, e.g. atoms attending over
// residues) and interleave over the streamed rows; their partial results meet in shared memory in a fixed order.
template <int VEC>
__device__ __forceinline__ void ld_vec(const float* __restrict__ p, float (&x)[VEC]) {
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
__device__ __forceinline__ void st_vec(float* __restrict__ p, const float (&x)[VEC]) {
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
// dot product of one head: partial over this lane's VEC columns, finished over the LPH lanes of the head (all get it)
template <int VEC, int LPH>
__device__ __forceinline__ float head_dot(const float (&a)[VEC], const float (&b)[VEC]) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) s = fmaf(a[i], b[i], s);
#pragma unroll
    for (int o = 1; o < LPH; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}

template <int WPR, int N>
__device__ __forceinline__ void warps_sum(float (&x)[N], float* sm, int w, int lane) {
    if constexpr (WPR > 1) {
#pragma unroll
        for (int i = 0; i < N; ++i) sm[(w * N + i) * 32 + lane] = x[i];
        __syncthreads();
#pragma unroll
        for (int i = 0; i < N; ++i) {
            float t = 0.f;
#pragma unroll
            for (int ww = 0; ww < WPR; ++ww) t += sm[(ww * N + i) * 32 + lane];
            x[i] = t;
        }
        __syncthreads();
    }
}

template <int H, int VEC, int WPR>
__global__ void __launch_bounds__(256) attn_fwd_kernel(const AttnArgs a) {
    constexpr int LPH = 32 / H, E = 32 * VEC;
    __shared__ float sm[WPR > 1 ? WPR * (VEC > 2 ? VEC : 2) * 32 : 1];
    const int lane = threadIdx.x & 31;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long row = gw / WPR;
    const int w = (int)(gw % WPR);
    const long long rows = a.Nq + (a.q_fill ? a.B : 0);
    if (row >= rows) return;                                 // block-uniform when WPR = warps per block
    const bool virt = row >= a.Nq;
    const long long g = virt ? row - a.Nq : a.qbatch[row];
    const long long k0 = a.kptr[g], k1 = a.kptr[g + 1];
    float q[VEC];
    ld_vec<VEC>((virt ? a.q_fill : a.q + row * E) + lane * VEC, q);
#pragma unroll
    for (int i = 0; i < VEC; ++i) q[i] *= a.scale;
    // pass 1: soft-max statistics (running max / sum), then merged over the warps of the row
    float m = -INFINITY, l = 0.f;
#pragma unroll 4
    for (long long j = k0 + w; j < k1; j += WPR) {
        float kk[VEC];
        ld_vec<VEC>(a.k + j * E + lane * VEC, kk);
        const float s = head_dot<VEC, LPH>(q, kk);
        const float mn = fmaxf(m, s);
        l = l * __expf(m - mn) + __expf(s - mn);
        m = mn;
    }
    if constexpr (WPR > 1) {
        sm[(w * 2) * 32 + lane] = m; sm[(w * 2 + 1) * 32 + lane] = l;
        __syncthreads();
        float mm = -INFINITY, ll = 0.f;
#pragma unroll
        for (int ww = 0; ww < WPR; ++ww) mm = fmaxf(mm, sm[(ww * 2) * 32 + lane]);
#pragma unroll
        for (int ww = 0; ww < WPR; ++ww) {
            const float mw = sm[(ww * 2) * 32 + lane];
            ll += mw == -INFINITY ? 0.f : sm[(ww * 2 + 1) * 32 + lane] * __expf(mw - mm);
        }
        __syncthreads();
        m = mm; l = ll;
    }
    const float inv = l > 0.f ? 1.f / l : 0.f;
    // pass 2: probabilities, output, head-averaged map
    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
    float* wrow = nullptr;
    if (virt) { if (a.w_fill) wrow = a.w_fill + g * a.lk_max; }
    else if (a.weights) wrow = a.weights + (g * a.lq_max + (row - a.qptr[g])) * a.lk_max;
#pragma unroll 4
    for (long long j = k0 + w; j < k1; j += WPR) {
        float kk[VEC], vv[VEC];
        ld_vec<VEC>(a.k + j * E + lane * VEC, kk);
        ld_vec<VEC>(a.v + j * E + lane * VEC, vv);
        const float p = __expf(head_dot<VEC, LPH>(q, kk) - m) * inv;
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = fmaf(p, vv[i], acc[i]);
        if (wrow) {                                          // uniform per warp
            float ws = p;
#pragma unroll
            for (int o = LPH; o < 32; o <<= 1) ws += __shfl_xor_sync(0xffffffffu, ws, o);
            if (lane == 0) wrow[j - k0] = ws * (1.f / (float)H);
        }
    }
    if (virt) return;
    warps_sum<WPR, VEC>(acc, sm, w, lane);
    if (w == 0) {
        st_vec<VEC>(a.out + row * E + lane * VEC, acc);
        if (lane % LPH == 0) reinterpret_cast<float2*>(a.stats)[row * H + lane / LPH] = make_float2(m, l);
    }
}

// dq[i,h] = scale * sum_j ds_ij k[j,h],  ds_ij = p_ij (dO_i . v_j - dO_i . O_i);  also writes dsum[i,h] = dO_i . O_i
template <int H, int VEC, int WPR>
__global__ void __launch_bounds__(256) attn_bwd_q_kernel(const AttnArgs a) {
    constexpr int LPH = 32 / H, E = 32 * VEC;
    __shared__ float sm[WPR > 1 ? WPR * VEC * 32 : 1];
    const int lane = threadIdx.x & 31, h = lane / LPH;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long row = gw / WPR;
    const int w = (int)(gw % WPR);
    if (row >= a.Nq) return;
    const long long g = a.qbatch[row];
    const long long k0 = a.kptr[g], k1 = a.kptr[g + 1];
    float q[VEC], go[VEC], oo[VEC];
    ld_vec<VEC>(a.q + row * E + lane * VEC, q);
    ld_vec<VEC>(a.d_out + row * E + lane * VEC, go);
    ld_vec<VEC>(a.o + row * E + lane * VEC, oo);
#pragma unroll
    for (int i = 0; i < VEC; ++i) q[i] *= a.scale;
    const float dsum = head_dot<VEC, LPH>(go, oo);
    const float2 st = reinterpret_cast<const float2*>(a.stats)[row * H + h];
    const float inv = st.y > 0.f ? 1.f / st.y : 0.f;
    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
#pragma unroll 4
    for (long long j = k0 + w; j < k1; j += WPR) {
        float kk[VEC], vv[VEC];
        ld_vec<VEC>(a.k + j * E + lane * VEC, kk);
        ld_vec<VEC>(a.v + j * E + lane * VEC, vv);
        const float p = __expf(head_dot<VEC, LPH>(q, kk) - st.x) * inv;
        const float ds = p * (head_dot<VEC, LPH>(go, vv) - dsum);
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = fmaf(ds, kk[i], acc[i]);
    }
    warps_sum<WPR, VEC>(acc, sm, w, lane);
    if (w == 0) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] *= a.scale;
        st_vec<VEC>(a.dq + row * E + lane * VEC, acc);
        if (lane % LPH == 0) a.dsum[row * H + h] = dsum;
    }
}

// dk[j,h] = scale * sum_i ds_ij q[i,h],  dv[j,h] = sum_i p_ij dO[i,h]      over the queries i of the key's graph
template <int H, int VEC, int WPR>
__global__ void __launch_bounds__(256) attn_bwd_kv_kernel(const AttnArgs a) {
    constexpr int LPH = 32 / H, E = 32 * VEC;
    __shared__ float sm[WPR > 1 ? WPR * 2 * VEC * 32 : 1];
    const int lane = threadIdx.x & 31, h = lane / LPH;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long row = gw / WPR;
    const int w = (int)(gw % WPR);
    if (row >= a.Nk) return;
    const long long g = a.kbatch[row];
    const long long q0 = a.qptr[g], q1 = a.qptr[g + 1];
    float kk[VEC], vv[VEC], dkv[2 * VEC];
    ld_vec<VEC>(a.k + row * E + lane * VEC, kk);
    ld_vec<VEC>(a.v + row * E + lane * VEC, vv);
#pragma unroll
    for (int i = 0; i < VEC; ++i) kk[i] *= a.scale;
#pragma unroll
    for (int i = 0; i < 2 * VEC; ++i) dkv[i] = 0.f;
#pragma unroll 4
    for (long long i = q0 + w; i < q1; i += WPR) {
        float q[VEC], go[VEC];
        ld_vec<VEC>(a.q + i * E + lane * VEC, q);
        ld_vec<VEC>(a.d_out + i * E + lane * VEC, go);
        const float2 st = __ldg(reinterpret_cast<const float2*>(a.stats) + i * H + h);
        const float dsum = __ldg(a.dsum + i * H + h);
        const float p = st.y > 0.f ? __expf(head_dot<VEC, LPH>(q, kk) - st.x) / st.y : 0.f;
        const float ds = p * (head_dot<VEC, LPH>(go, vv) - dsum);
#pragma unroll
        for (int d = 0; d < VEC; ++d) { dkv[VEC + d] = fmaf(p, go[d], dkv[VEC + d]); dkv[d] = fmaf(ds, q[d], dkv[d]); }
    }
    warps_sum<WPR, 2 * VEC>(dkv, sm, w, lane);
    if (w == 0) {
        float dk[VEC], dv[VEC];
#pragma unroll
        for (int d = 0; d < VEC; ++d) { dk[d] = dkv[d] * a.scale; dv[d] = dkv[VEC + d]; }
        st_vec<VEC>(a.dk + row * E + lane * VEC, dk);
        st_vec<VEC>(a.dv + row * E + lane * VEC, dv);
    }
}

template <int H, int VEC, int WPR>
static int attn_launch_w(const AttnArgs& a, int which, long long rows, cudaStream_t st) {
    const unsigned grid = (unsigned)cdiv64(rows * WPR * 32, 256);
    if (which == 0) attn_fwd_kernel<H, VEC, WPR><<<grid, 256, 0, st>>>(a);
    else if (which == 1) attn_bwd_q_kernel<H, VEC, WPR><<<grid, 256, 0, st>>>(a);
    else attn_bwd_kv_kernel<H, VEC, WPR><<<grid, 256, 0, st>>>(a);
    CGVP_LAUNCH_CHECK("attention kernel");
    return 0;
}

template <int H, int VEC>
static int attn_launch(const AttnArgs& a, int which, cudaStream_t st) {
    const long long rows = which == 0 ? a.Nq + (a.q_fill ? a.B : 0) : (which == 1 ? a.Nq : a.Nk);
    if (rows <= 0) return 0;
    // few rows, long streams (e.g. ~10^3 atoms over ~10^3 residues each): spread one row over the 8 warps of a CTA
    const long long stream_len = (which == 2 ? a.Nq : a.Nk) / (a.B > 0 ? a.B : 1);
    const bool wide = rows * 32 < (long long)cgvp_num_sms() * 1024 && stream_len >= 64;
    return wide ? attn_launch_w<H, VEC, 8>(a, which, rows, st) : attn_launch_w<H, VEC, 1>(a, which, rows, st);
}

// supported: heads in {1, 2, 4, 8, 16, 32}, embedding E = heads * head_dim in {32, 64, 128, 256}
static int attn_dispatch(const AttnArgs& a, int H, int D, int which, cudaStream_t st) {
    const int E = H * D;
#define CGVP_ATTN_CASE(h, vec) if (H == h && E == 32 * vec) return attn_launch<h, vec>(a, which, st);
#define CGVP_ATTN_HEADS(vec) CGVP_ATTN_CASE(1, vec) CGVP_ATTN_CASE(2, vec) CGVP_ATTN_CASE(4, vec) CGVP_ATTN_CASE(8, vec) \
                             CGVP_ATTN_CASE(16, vec) CGVP_ATTN_CASE(32, vec)
    CGVP_ATTN_HEADS(1) CGVP_ATTN_HEADS(2) CGVP_ATTN_HEADS(4) CGVP_ATTN_HEADS(8)
#undef CGVP_ATTN_HEADS
#undef CGVP_ATTN_CASE
    cgvp_set_error("attention: no kernel for %d heads of width %d", H, D);
    return -2;
}

extern "C" int32_t cgvp_attn_supported(int32_t num_heads, int32_t head_dim) {
    const int e = num_heads * head_dim;
    const bool heads_ok = num_heads == 1 || num_heads == 2 || num_heads == 4 || num_heads == 8 || num_heads == 16 || num_heads == 32;
    return heads_ok && (e == 32 || e == 64 || e == 128 || e == 256) ? 1 : 0;
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int32_t cgvp_attn_fwd(const float* q, const float* k, const float* v, const int64_t* qptr, const int64_t* kptr,
                                 const int64_t* qbatch, int64_t num_graphs, int64_t num_q, int64_t num_k, int32_t num_heads,
                                 int32_t head_dim, float scale, const float* q_fill, int32_t lq_max, int32_t lk_max, float* out,
                                 float* stats, float* weights, float* w_fill, void* stream) {
    CGVP_REQUIRE(num_graphs >= 0 && num_q >= 0 && num_k >= 0, "attn_fwd: bad sizes");
    if (num_q == 0) return 0;
    CGVP_REQUIRE(q && k && v && qptr && kptr && qbatch && out && stats, "attn_fwd: null argument");
    CGVP_REQUIRE(al16(q) && al16(k) && al16(v) && al16(out) && al16(stats) && al16(q_fill), "attn_fwd: buffers must be 16-byte aligned");
    CGVP_REQUIRE(!weights || (lq_max > 0 && lk_max > 0), "attn_fwd: attention map without its padded shape");
    AttnArgs a;
    memset(&a, 0, sizeof(a));
    a.q = q; a.k = k; a.v = v; a.qptr = qptr; a.kptr = kptr; a.qbatch = qbatch; a.Nq = num_q; a.Nk = num_k; a.B = num_graphs;
    a.scale = scale; a.lq_max = lq_max; a.lk_max = lk_max; a.q_fill = (weights && w_fill) ? q_fill : nullptr;
    a.out = out; a.stats = stats; a.weights = weights; a.w_fill = w_fill;
    return attn_dispatch(a, num_heads, head_dim, 0, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int32_t cgvp_attn_bwd(const float* q, const float* k, const float* v, const float* out, const float* stats,
                                 const float* d_out, const int64_t* qptr, const int64_t* kptr, const int64_t* qbatch,
                                 const int64_t* kbatch, int64_t num_graphs, int64_t num_q, int64_t num_k, int32_t num_heads,
                                 int32_t head_dim, float scale, float* dsum, float* dq, float* dk, float* dv, void* stream) {
    CGVP_REQUIRE(num_graphs >= 0 && num_q >= 0 && num_k >= 0, "attn_bwd: bad sizes");
    CGVP_REQUIRE(q && k && v && out && stats && d_out && qptr && kptr && qbatch && kbatch && dsum && dq && dk && dv,
                 "attn_bwd: null argument");
    CGVP_REQUIRE(al16(q) && al16(k) && al16(v) && al16(out) && al16(stats) && al16(d_out) && al16(dq) && al16(dk) && al16(dv),
                 "attn_bwd: buffers must be 16-byte aligned");
    AttnArgs a;
    memset(&a, 0, sizeof(a));
    a.q = q; a.k = k; a.v = v; a.o = out; a.stats = const_cast<float*>(stats); a.d_out = d_out; a.qptr = qptr; a.kptr = kptr; a.qbatch = qbatch;
    a.kbatch = kbatch; a.Nq = num_q; a.Nk = num_k; a.B = num_graphs; a.scale = scale; a.dsum = dsum; a.dq = dq; a.dk = dk; a.dv = dv;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc = attn_dispatch(a, num_heads, head_dim, 1, st);
    if (rc) return rc;
    return attn_dispatch(a, num_heads, head_dim, 2, st);
}

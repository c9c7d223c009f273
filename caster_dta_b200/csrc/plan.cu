// Graph plan (dst-sorted and src-sorted CSR views of edge_index) and the two stand-alone HBM-bound primitives:
// message-input gather and deterministic segmented reduction.  Replaces the index handling of PyG
// MessagePassing.propagate as used by GVPConv (models/gvp_layers.py:298-300).
#include <cub/device/device_radix_sort.cuh>

#include "cgvp_common.cuh"

// ---- plan ---------------------------------------------------------------------------------------------------------
__global__ void plan_prepare_kernel(const int64_t* __restrict__ edge_index, int64_t E, int* __restrict__ key_dst,
                                    int* __restrict__ iota) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    key_dst[e] = (int)edge_index[E + e];
    iota[e] = (int)e;
}

__global__ void plan_gather_src_kernel(const int64_t* __restrict__ edge_index, int64_t E, const int* __restrict__ perm,
                                       int* __restrict__ src_sorted) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= E) return;
    src_sorted[p] = (int)edge_index[perm[p]];
}

// rowptr[n] = first position whose (sorted) key is >= n
__global__ void plan_rowptr_kernel(const int* __restrict__ sorted_keys, int64_t E, int64_t N, int* __restrict__ rowptr) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n > N) return;
    int64_t lo = 0, hi = E;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (sorted_keys[mid] < n) lo = mid + 1; else hi = mid;
    }
    rowptr[n] = (int)lo;
}

static int bits_for(int64_t n) {
    int b = 1;
    while (((int64_t)1 << b) < n && b < 31) ++b;
    return b;
}

static size_t cub_sort_bytes(int64_t E, int64_t N) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int*)nullptr, (int*)nullptr, (const int*)nullptr,
                                    (int*)nullptr, (int)E, 0, bits_for(N));
    return bytes;
}

extern "C" int64_t cgvp_plan_workspace_bytes(int64_t num_edges, int64_t num_nodes) {
    if (num_edges < 0 || num_nodes < 0 || num_edges >= ((int64_t)1 << 31) || num_nodes >= ((int64_t)1 << 31)) return -1;
    const int64_t e4 = align_up(num_edges * 4, 256);
    return 3 * e4 + (int64_t)align_up((int64_t)cub_sort_bytes(num_edges, num_nodes), 256) + 256;
}

extern "C" int32_t cgvp_plan_build(const int64_t* edge_index, const CgvpPlan* plan, void* ws, int64_t ws_bytes,
                                   cgvp_stream_t stream) {
    CGVP_REQUIRE(plan, "plan_build: null plan");
    const int64_t E = plan->num_edges, N = plan->num_nodes;
    CGVP_REQUIRE(E >= 0 && N >= 0 && E < ((int64_t)1 << 31) && N < ((int64_t)1 << 31), "plan_build: sizes out of range");
    CGVP_REQUIRE(plan->rowptr && plan->srowptr, "plan_build: null rowptr");
    cudaStream_t st = (cudaStream_t)stream;
    const int T = 256;
    if (E > 0) {
        CGVP_REQUIRE(edge_index && plan->perm && plan->src && plan->dst && plan->sperm, "plan_build: null buffer");
        const int64_t need = cgvp_plan_workspace_bytes(E, N);
        CGVP_REQUIRE(ws && ws_bytes >= need, "plan_build: workspace too small (%lld < %lld)", (long long)ws_bytes,
                     (long long)need);
        const int64_t e4 = align_up(E * 4, 256);
        char* base = reinterpret_cast<char*>(ws);
        int* key = reinterpret_cast<int*>(base);
        int* iota = reinterpret_cast<int*>(base + e4);
        int* skey = reinterpret_cast<int*>(base + 2 * e4);
        void* cub_tmp = base + 3 * e4;
        size_t cub_bytes = cub_sort_bytes(E, N);
        const int grid = (int)cdiv64(E, T);
        plan_prepare_kernel<<<grid, T, 0, st>>>(edge_index, E, key, iota);
        CGVP_LAUNCH_CHECK("plan_prepare_kernel");
        // stable LSD radix sort by target: ties keep the original edge order, so the summation order is reproducible
        CGVP_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, (const int*)key, plan->dst, (const int*)iota,
                                                  plan->perm, (int)E, 0, bits_for(N), st));
        plan_gather_src_kernel<<<grid, T, 0, st>>>(edge_index, E, plan->perm, plan->src);
        CGVP_LAUNCH_CHECK("plan_gather_src_kernel");
        // second view: dst-sorted positions ordered by source
        CGVP_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, (const int*)plan->src, skey, (const int*)iota,
                                                  plan->sperm, (int)E, 0, bits_for(N), st));
        plan_rowptr_kernel<<<(int)cdiv64(N + 1, T), T, 0, st>>>(plan->dst, E, N, plan->rowptr);
        CGVP_LAUNCH_CHECK("plan_rowptr_kernel");
        plan_rowptr_kernel<<<(int)cdiv64(N + 1, T), T, 0, st>>>(skey, E, N, plan->srowptr);
        CGVP_LAUNCH_CHECK("plan_rowptr_kernel");
    } else {
        CGVP_CUDA(cudaMemsetAsync(plan->rowptr, 0, (size_t)(N + 1) * 4, st));
        CGVP_CUDA(cudaMemsetAsync(plan->srowptr, 0, (size_t)(N + 1) * 4, st));
    }
    return 0;
}

// ---- gather of the message input (gvp_layers.py:303-306) ----------------------------------------------------------
// A CTA of (columns per edge) x (edges per pass) threads: every thread owns ONE output column (a float4 of the scalar
// row, or a float of the vector row) for its whole life, so which source tensor it reads, and at which offset, is
// decided once; the edge loop is index load -> row load -> coalesced store with no division.
template <class T>   // T = float4 (16-byte columns) or float
__global__ void __launch_bounds__(256) gather_rows_kernel(const int64_t* __restrict__ ei, int64_t E, int w_node, int w_edge,
                                                          const T* __restrict__ node, const T* __restrict__ edge,
                                                          T* __restrict__ out) {
    const int w = 2 * w_node + w_edge;              // columns per output row
    const int epp = blockDim.x / w;                 // edges per pass of this CTA
    const int c = threadIdx.x % w, de = threadIdx.x / w;
    if (de >= epp) return;
    // 0: source row of `node`, 1: the edge's own row, 2: target row of `node`
    const int kind = c < w_node ? 0 : (c < w_node + w_edge ? 1 : 2);
    const int off = kind == 0 ? c : (kind == 1 ? c - w_node : c - w_node - w_edge);
    const int64_t stride = (int64_t)gridDim.x * epp;
    constexpr int U = 2;                            // edges in flight per thread: index loads, then row loads, then stores
    int64_t e = (int64_t)blockIdx.x * epp + de;
    for (; e + (U - 1) * stride < E; e += U * stride) {
        const T* p[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t eu = e + u * stride;
            p[u] = kind == 1 ? edge + eu * w_edge + off : node + __ldg(ei + (kind == 0 ? eu : E + eu)) * w_node + off;
        }
        T v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = kind == 1 ? __ldcs(p[u]) : __ldg(p[u]);     // edge rows stream, node rows are reused
#pragma unroll
        for (int u = 0; u < U; ++u) __stcs(out + (e + u * stride) * w + c, v[u]);       // the output is never re-read here
    }
    for (; e < E; e += stride) {
        T v;
        if (kind == 1) v = __ldg(edge + e * w_edge + off);
        else v = __ldg(node + __ldg(ei + (kind == 0 ? e : E + e)) * w_node + off);
        out[e * w + c] = v;
    }
}

// Rows whose width is not a multiple of 16 bytes (the vector part: 3 * (2 nv + ev) floats): a CTA assembles
// GV_EDGES message rows in shared memory (float4 loads of the node rows where they are 16-byte aligned, scalars for
// the rest) and writes the assembled chunk -- contiguous in the output -- with coalesced float4 stores.
#define GV_EDGES 128
__global__ void __launch_bounds__(256) gather_staged_kernel(const int64_t* __restrict__ ei, int64_t E, int w_node, int w_edge,
                                                            const float* __restrict__ node, const float* __restrict__ edge,
                                                            float* __restrict__ out, int node_vec4) {
    extern __shared__ __align__(16) float gbuf[];
    const int w = 2 * w_node + w_edge;
    const int nq = node_vec4 ? w_node / 4 : w_node;        // load units per node row
    const int units = 2 * nq + w_edge;                     // per edge
    const int64_t nchunks = cdiv64(E, GV_EDGES);
    for (int64_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const int64_t e0 = ch * GV_EDGES;
        const int ne = (int)(E - e0 < GV_EDGES ? E - e0 : GV_EDGES);
        for (int i = threadIdx.x; i < ne * units; i += blockDim.x) {
            const int le = i / units, u = i - le * units;
            const int64_t e = e0 + le;
            float* dst = gbuf + le * w;
            if (u < 2 * nq) {
                const int side = u >= nq, j = side ? u - nq : u;
                const int64_t n = __ldg(ei + (side ? E + e : e));
                float* d = dst + (side ? w_node + w_edge : 0);
                if (node_vec4) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(node + n * w_node) + j);
                    d[4 * j] = t.x; d[4 * j + 1] = t.y; d[4 * j + 2] = t.z; d[4 * j + 3] = t.w;
                } else {
                    d[j] = __ldg(node + n * w_node + j);
                }
            } else {
                const int j = u - 2 * nq;
                dst[w_node + j] = __ldg(edge + e * w_edge + j);
            }
        }
        __syncthreads();
        const int total = ne * w;                          // floats of this chunk; its start e0 * w * 4 bytes is 16-byte aligned
        float* o = out + e0 * w;
        const int t4 = total >> 2;
        for (int i = threadIdx.x; i < t4; i += blockDim.x)
            __stcs(reinterpret_cast<float4*>(o) + i, reinterpret_cast<const float4*>(gbuf)[i]);
        for (int i = (t4 << 2) + threadIdx.x; i < total; i += blockDim.x) o[i] = gbuf[i];
        __syncthreads();
    }
}

template <class T>
static int launch_gather_rows(const int64_t* ei, int64_t E, int w_node, int w_edge, const T* node, const T* edge, T* out,
                              int sms, cudaStream_t st) {
    const int w = 2 * w_node + w_edge;
    CGVP_REQUIRE(w <= 1024, "gather: message row of %d columns is too wide", w);
    const int epp = w <= 256 ? 256 / w : 1;
    const int threads = w * epp;
    const int64_t want = cdiv64(E, (int64_t)epp * 8);     // >= 8 edges per thread
    const int grid = (int)(want < (int64_t)sms * 8 ? (want > 0 ? want : 1) : (int64_t)sms * 8);
    gather_rows_kernel<T><<<grid, threads, 0, st>>>(ei, E, w_node, w_edge, node, edge, out);
    CGVP_LAUNCH_CHECK("gather_rows_kernel");
    return 0;
}

extern "C" int32_t cgvp_gather_message_input(const int64_t* edge_index, int64_t num_edges, int32_t ns, int32_t nv,
                                             int32_t es, int32_t ev, const float* s, const float* v, const float* e_s,
                                             const float* e_v, float* ms, float* mv, cgvp_stream_t stream) {
    CGVP_REQUIRE(num_edges >= 0 && ns >= 0 && nv >= 0 && es >= 0 && ev >= 0, "gather: bad sizes");
    if (num_edges == 0) return 0;
    CGVP_REQUIRE(edge_index, "gather: null edge_index");
    cudaStream_t st = (cudaStream_t)stream;
    const int sms = cgvp_num_sms() > 0 ? cgvp_num_sms() : 148;
    cgvp_prof_begin(CGVP_K_GATHER, st);
    if (2 * ns + es > 0) {
        CGVP_REQUIRE(ms && (ns == 0 || s) && (es == 0 || e_s), "gather: null scalar buffer");
        const bool vec = ns % 4 == 0 && es % 4 == 0 && (((uintptr_t)s | (uintptr_t)e_s | (uintptr_t)ms) & 15) == 0;
        int rc;
        if (vec) rc = launch_gather_rows<float4>(edge_index, num_edges, ns / 4, es / 4, reinterpret_cast<const float4*>(s),
                                                 reinterpret_cast<const float4*>(e_s), reinterpret_cast<float4*>(ms), sms, st);
        else rc = launch_gather_rows<float>(edge_index, num_edges, ns, es, s, e_s, ms, sms, st);
        if (rc) return rc;
    }
    if (2 * nv + ev > 0) {
        CGVP_REQUIRE(mv && (nv == 0 || v) && (ev == 0 || e_v), "gather: null vector buffer");
        const bool vec = (3 * nv) % 4 == 0 && (3 * ev) % 4 == 0 && (((uintptr_t)v | (uintptr_t)e_v | (uintptr_t)mv) & 15) == 0;
        int rc;
        if (vec) rc = launch_gather_rows<float4>(edge_index, num_edges, 3 * nv / 4, 3 * ev / 4, reinterpret_cast<const float4*>(v),
                                                 reinterpret_cast<const float4*>(e_v), reinterpret_cast<float4*>(mv), sms, st);
        else if ((((uintptr_t)mv) & 15) == 0 && (size_t)GV_EDGES * 3 * (2 * nv + ev) * 4 <= 48 * 1024) {
            const int node_vec4 = (3 * nv) % 4 == 0 && (((uintptr_t)v) & 15) == 0;
            const int64_t nch = cdiv64(num_edges, GV_EDGES);
            const int grid = (int)(nch < (int64_t)sms * 8 ? nch : (int64_t)sms * 8);
            gather_staged_kernel<<<grid, 256, (size_t)GV_EDGES * 3 * (2 * nv + ev) * 4, st>>>(edge_index, num_edges, 3 * nv, 3 * ev, v, e_v,
                                                                                           mv, node_vec4);
            CGVP_LAUNCH_CHECK("gather_staged_kernel");
            rc = 0;
        } else rc = launch_gather_rows<float>(edge_index, num_edges, 3 * nv, 3 * ev, v, e_v, mv, sms, st);
        if (rc) return rc;
    }
    cgvp_prof_end(CGVP_K_GATHER, st);
    return 0;
}

// ---- deterministic segmented reduction -------------------------------------------------------------------------------
// out[n][c] (+)= scale(n) * sum_{p in [rowptr[n], rowptr[n+1])} rows[index[p]][c], summed in position order.
// The output may be split into two tensors (widths wa | wb) so merged message rows can land in (s, V) pairs.
template <int V>
__global__ void __launch_bounds__(256) segment_reduce_kernel(const float* __restrict__ rows, int width,
                                                             const int* __restrict__ rowptr, const int* __restrict__ index,
                                                             int64_t N, int aggr, int beta, float* __restrict__ out_a,
                                                             int wa, float* __restrict__ out_b, int wb) {
    const int wv = width / V;
    const int64_t total = N * wv;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = i / wv;
        const int c = (int)(i - n * wv) * V;
        const int p0 = rowptr[n], p1 = rowptr[n + 1];
        float acc[V];
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] = 0.f;
        for (int p = p0; p < p1; ++p) {
            const int64_t r = index ? index[p] : p;
            if (V == 4) {
                const float4 x = __ldg(reinterpret_cast<const float4*>(rows + r * width + c));
                acc[0] += x.x; acc[1 % V] += x.y; acc[2 % V] += x.z; acc[3 % V] += x.w;
            } else {
                acc[0] += __ldg(rows + r * width + c);
            }
        }
        const float scale = (aggr == CGVP_AGGR_MEAN) ? 1.f / (float)max(p1 - p0, 1) : 1.f;
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const int cc = c + k;
            float* dst = cc < wa ? out_a + n * wa + cc : out_b + n * wb + (cc - wa);
            const float val = acc[k] * scale;
            *dst = beta ? *dst + val : val;
        }
    }
}

int cgvp_segment_reduce_split(const float* rows, int width, const int* rowptr, const int* index, int64_t N, int aggr,
                              int beta, float* out_a, int wa, float* out_b, int wb, cudaStream_t st) {
    if (N == 0 || width == 0) return 0;
    const int sms = cgvp_num_sms() > 0 ? cgvp_num_sms() : 148;
    const bool vec = width % 4 == 0 && ((uintptr_t)rows & 15) == 0;
    const int64_t total = N * (width / (vec ? 4 : 1));
    const int grid = (int)(cdiv64(total, 256) < (int64_t)sms * 16 ? cdiv64(total, 256) : (int64_t)sms * 16);
    cgvp_prof_begin(CGVP_K_SEGMENT_REDUCE, st);
    if (vec) segment_reduce_kernel<4><<<grid, 256, 0, st>>>(rows, width, rowptr, index, N, aggr, beta, out_a, wa, out_b, wb);
    else segment_reduce_kernel<1><<<grid, 256, 0, st>>>(rows, width, rowptr, index, N, aggr, beta, out_a, wa, out_b, wb);
    cgvp_prof_end(CGVP_K_SEGMENT_REDUCE, st);
    CGVP_LAUNCH_CHECK("segment_reduce_kernel");
    return 0;
}

extern "C" int32_t cgvp_segment_reduce(const float* rows, int32_t width, const int32_t* rowptr, const int32_t* index,
                                       int64_t num_nodes, int32_t aggr, int32_t beta, float* out, cgvp_stream_t stream) {
    CGVP_REQUIRE(width >= 0 && num_nodes >= 0, "segment_reduce: bad sizes");
    if (num_nodes == 0 || width == 0) return 0;
    CGVP_REQUIRE(rows && rowptr && out, "segment_reduce: null buffer");
    return cgvp_segment_reduce_split(rows, width, rowptr, index, num_nodes, aggr, beta, out, width, nullptr, 0,
                                     (cudaStream_t)stream);
}

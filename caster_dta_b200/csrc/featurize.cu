// Residue-graph featurizer for a batch of proteins: radius / kNN / proportional edge selection on C-alpha distances,
// 16 Gaussian RBFs, 16 sin/cos sequence-offset encodings and the unit direction vector, written directly in COO
// order (src ascending, dst ascending).  Replaces utils/create_protein_features.py:201-357 followed by
// utils/create_graphs.py:6-62 without ever materialising the dense n x n x 35 fp64 array.
//
// Bit-exactness: the reference thresholds fp64 distances from scipy `pdist` on fp32 coordinates, i.e.
// sqrt(((dx*dx) + dy*dy) + dz*dz) with every operation rounded separately.  The __d*_rn intrinsics below are never
// contracted into FMAs, so the distance -- and therefore the edge set and edge_index -- is bit-identical.  The
// direction vector follows numpy's fp32 arithmetic the same way.  RBF / sin / cos go through CUDA's fp64 libm
// (<= 1-2 ulp in fp64) before the fp32 cast, so those features agree to <= 1 ulp(fp32).
#include <cub/device/device_scan.cuh>

#include "cgvp_common.cuh"

#define FEAT_THREADS 128
#define FEAT_WARPS (FEAT_THREADS / 32)

// np.linspace(0., 20., 16) and np.exp(2*np.arange(8) * -(np.log(10000.0)/8)), printed with float.hex() from numpy 2.3
__constant__ double c_rbf_mu[16] = {0x0.0p+0, 0x1.5555555555555p+0, 0x1.5555555555555p+1, 0x1.0000000000000p+2,
                                    0x1.5555555555555p+2, 0x1.aaaaaaaaaaaaap+2, 0x1.0000000000000p+3, 0x1.2aaaaaaaaaaaap+3,
                                    0x1.5555555555555p+3, 0x1.8000000000000p+3, 0x1.aaaaaaaaaaaaap+3, 0x1.d555555555555p+3,
                                    0x1.0000000000000p+4, 0x1.1555555555555p+4, 0x1.2aaaaaaaaaaaap+4, 0x1.4000000000000p+4};
__constant__ double c_pe_freq[8] = {0x1.0000000000000p+0, 0x1.9999999999998p-4, 0x1.47ae147ae1478p-7, 0x1.0624dd2f1a9f9p-10,
                                    0x1.a36e2eb1c4326p-14, 0x1.4f8b588e368e5p-17, 0x1.0c6f7a0b5ed87p-20, 0x1.ad7f29abcaf44p-24};

struct FeatK {
    const float* ca;
    const int64_t* ptr;
    int64_t B, N;
    double thresh;
    int type, keep_self;
    int* deg;
    const int64_t* row_offsets;
    int64_t *ei, E;
    float *es, *ev;
    const float* pe;      // [2 * max_len - 1][16] positional encodings by (dst - src) + max_len - 1, or NULL
    int64_t max_len;
    int cache_rows;       // shared-memory distance cache (doubles per warp), 0 = recompute
};

__device__ __forceinline__ double ca_distance(const float* __restrict__ ca, int64_t i, int64_t j) {
    const double dx = (double)ca[3 * i] - (double)ca[3 * j];
    const double dy = (double)ca[3 * i + 1] - (double)ca[3 * j + 1];
    const double dz = (double)ca[3 * i + 2] - (double)ca[3 * j + 2];
    return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
}

__device__ __forceinline__ int64_t find_protein(const int64_t* __restrict__ ptr, int64_t B, int64_t i) {
    int64_t lo = 0, hi = B;   // largest b with ptr[b] <= i
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (ptr[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ int knn_k(const FeatK& K, int64_t n) {
    if (K.type == 2) return (int)ceil(K.thresh * (double)n);   // int(np.ceil(edge_thresh * n_residues))
    return (int)K.thresh;                                      // int(edge_thresh)
}

// ---- phase 1: edges per source residue ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(FEAT_THREADS) feat_count_kernel(const FeatK K) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < K.N; i += nwarps) {
        const int64_t b = find_protein(K.ptr, K.B, i);
        const int64_t lo = K.ptr[b], hi = K.ptr[b + 1], n = hi - lo;
        int cnt;
        if (K.type == 0) {
            cnt = 0;
            for (int64_t j0 = lo; j0 < hi; j0 += 32) {
                const int64_t j = j0 + lane;
                bool keep = false;
                if (j < hi && (K.keep_self || j != i)) keep = ca_distance(K.ca, i, j) <= K.thresh;
                cnt += __popc(__ballot_sync(0xffffffffu, keep));
            }
        } else {
            const int64_t valid = n - (K.keep_self ? 0 : 1);
            const int64_t k = knn_k(K, n);
            cnt = (int)(k < valid ? (k < 0 ? 0 : k) : valid);
        }
        if (lane == 0) K.deg[i] = cnt;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) K.deg[K.N] = 0;
}

// positional encodings depend on (dst - src) only: [cos(delta f_0..7) ; sin(delta f_0..7)], fp64 then cast  (:368-385)
__global__ void feat_pe_kernel(int64_t max_len, float* __restrict__ pe) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (2 * max_len - 1) * 16) return;
    const int64_t delta = i / 16 - (max_len - 1);
    const int f = (int)(i % 16);
    const double ang = (double)delta * c_pe_freq[f & 7];
    pe[i] = (float)(f < 8 ? cos(ang) : sin(ang));
}

// ---- phase 2: write edges and features -----------------------------------------------------------------------------------
// One LANE per edge: the 16 fp64 exponentials of an edge's RBF row are independent (ILP), and up to 32 edges of the
// row are in flight per warp.  (A lane-per-feature layout serialises the edges of a row behind the ~500-cycle latency
// of one fp64 exp.)
__device__ __forceinline__ void emit_edge_lane(const FeatK& K, int64_t e, int64_t i, int64_t j, double d) {
    float f[32];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const double t = (d - c_rbf_mu[k]) / 1.25;                           // (d - mu) / D_step            :237
        f[k] = (float)exp(-(t * t));
    }
    if (K.pe) {                                                              // tabulated with the same fp64 formula
        const float4* row = reinterpret_cast<const float4*>(K.pe + ((j - i) + K.max_len - 1) * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 t = __ldg(row + q);
            f[16 + 4 * q] = t.x; f[17 + 4 * q] = t.y; f[18 + 4 * q] = t.z; f[19 + 4 * q] = t.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const double ang = (double)(j - i) * c_pe_freq[k];               // (dst idx - src idx) * freq   :253,:382
            f[16 + k] = (float)cos(ang);
            f[24 + k] = (float)sin(ang);
        }
    }
    float4* es4 = reinterpret_cast<float4*>(K.es + e * 32);
#pragma unroll
    for (int q = 0; q < 8; ++q) es4[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
    const float x = __fsub_rn(K.ca[3 * i], K.ca[3 * j]), y = __fsub_rn(K.ca[3 * i + 1], K.ca[3 * j + 1]),
                z = __fsub_rn(K.ca[3 * i + 2], K.ca[3 * j + 2]);                                        // :244
    const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
    K.ev[e * 3] = nrm != 0.f ? __fdiv_rn(x, nrm) : 0.f;                                                  // :360-365
    K.ev[e * 3 + 1] = nrm != 0.f ? __fdiv_rn(y, nrm) : 0.f;
    K.ev[e * 3 + 2] = nrm != 0.f ? __fdiv_rn(z, nrm) : 0.f;
    K.ei[e] = i;
    K.ei[K.E + e] = j;
}

// per-warp list of kept (column, distance) pairs, flushed 32 edges at a time in column order
struct EdgeList {
    int64_t* j;       // [64]
    double* d;        // [64]
    int cnt;
};
__device__ __forceinline__ void list_flush(const FeatK& K, EdgeList& L, int64_t& e, int64_t i, int lane, int n) {
    if (lane < n) emit_edge_lane(K, e + lane, i, L.j[lane], L.d[lane]);
    e += n;
    __syncwarp();
    if (L.cnt > n) {                                       // move the tail (< 32 entries) to the front
        const bool mv = lane < L.cnt - n;
        int64_t tj = 0; double td = 0.0;
        if (mv) { tj = L.j[n + lane]; td = L.d[n + lane]; }
        __syncwarp();
        if (mv) { L.j[lane] = tj; L.d[lane] = td; }
    }
    L.cnt -= n;
    __syncwarp();
}
__device__ __forceinline__ void emit_chunk(const FeatK& K, EdgeList& L, unsigned mask, int64_t& e, int64_t i, int64_t j, double d, int lane) {
    if (mask >> lane & 1u) {
        const int pos = L.cnt + __popc(mask & ((1u << lane) - 1u));
        L.j[pos] = j; L.d[pos] = d;
    }
    L.cnt += __popc(mask);
    __syncwarp();
    if (L.cnt >= 32) list_flush(K, L, e, i, lane, 32);
}

__global__ void __launch_bounds__(FEAT_THREADS) feat_fill_kernel(const FeatK K) {
    __shared__ int hist[FEAT_WARPS][256];
    __shared__ int64_t list_j[FEAT_WARPS][64];
    __shared__ double list_d[FEAT_WARPS][64];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    EdgeList L;
    L.j = list_j[wib]; L.d = list_d[wib]; L.cnt = 0;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < K.N; i += nwarps) {
        const int64_t b = find_protein(K.ptr, K.B, i);
        const int64_t lo = K.ptr[b], hi = K.ptr[b + 1], n = hi - lo;
        int64_t e = K.row_offsets[i];
        const int64_t e_end = K.row_offsets[i + 1];
        if (e_end == e) continue;
        if (K.type == 0) {
            for (int64_t j0 = lo; j0 < hi; j0 += 32) {
                const int64_t j = j0 + lane;
                double d = 0.0;
                bool keep = false;
                if (j < hi && (K.keep_self || j != i)) { d = ca_distance(K.ca, i, j); keep = d <= K.thresh; }
                emit_chunk(K, L, __ballot_sync(0xffffffffu, keep), e, i, j, d, lane);
            }
            if (L.cnt > 0) list_flush(K, L, e, i, lane, L.cnt);
            continue;
        }
        // k nearest: radix-select the k-th smallest distance (fp64 bit pattern of a non-negative double is monotone),
        // then keep d < T and the first (k - #{d < T}) ties in column order  (np.argsort(...)[:, :k], :319)
        const int kk = (int)(e_end - e);
        // fp64 distances of this row, computed ONCE into the warp's shared-memory cache (excluded self = +inf)
        extern __shared__ __align__(16) unsigned long long dcache[];
        unsigned long long* drow = dcache + (size_t)wib * K.cache_rows;
        const bool cached = K.cache_rows >= n;
        if (cached) {
            for (int64_t j = lo + lane; j < hi; j += 32)
                drow[j - lo] = (K.keep_self || j != i) ? (unsigned long long)__double_as_longlong(ca_distance(K.ca, i, j))
                                                       : 0x7ff0000000000000ull;
            __syncwarp();
        }
        unsigned long long prefix = 0ull, pmask = 0ull;
        int remaining = kk;
        for (int shift = 56; shift >= 0; shift -= 8) {
            for (int q = lane; q < 256; q += 32) hist[wib][q] = 0;
            __syncwarp();
            for (int64_t j0 = lo; j0 < hi; j0 += 32) {
                const int64_t j = j0 + lane;
                if (j < hi && (K.keep_self || j != i)) {
                    const unsigned long long key = cached ? drow[j - lo]
                                                          : (unsigned long long)__double_as_longlong(ca_distance(K.ca, i, j));
                    if ((key & pmask) == prefix) atomicAdd(&hist[wib][(int)((key >> shift) & 255ull)], 1);
                }
            }
            __syncwarp();
            // each lane owns 8 consecutive bins; warp-exclusive scan of the lane totals
            int local[8], tot = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { local[q] = hist[wib][lane * 8 + q]; tot += local[q]; }
            int incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            int run = incl - tot, bin = -1, before = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (bin < 0 && run + local[q] >= remaining) { bin = lane * 8 + q; before = run; }
                run += local[q];
            }
            const unsigned found = __ballot_sync(0xffffffffu, bin >= 0);
            const int srcl = __ffs(found) - 1;
            bin = __shfl_sync(0xffffffffu, bin, srcl);
            before = __shfl_sync(0xffffffffu, before, srcl);
            remaining -= before;
            prefix |= (unsigned long long)bin << shift;
            pmask |= 255ull << shift;
            const int in_bin = hist[wib][bin];
            __syncwarp();
            // every key that shares the prefix is taken: the lower digits cannot change the selection
            if (in_bin == remaining) break;
        }
        // keep keys whose masked bits are below the prefix, and the first `remaining` keys (column order) that match it.
        // (fp64 bit patterns of non-negative doubles order like the doubles; after all 8 digits "match" means d == T, which
        //  reproduces np.argsort(...)[:, :k] with ties in column order, :319)
        int ties_left = remaining;
        for (int64_t j0 = lo; j0 < hi; j0 += 32) {
            const int64_t j = j0 + lane;
            double d = 0.0;
            bool lt = false, eq = false;
            if (j < hi && (K.keep_self || j != i)) {
                const unsigned long long key = cached ? drow[j - lo] : (unsigned long long)__double_as_longlong(ca_distance(K.ca, i, j));
                d = __longlong_as_double((long long)key);
                lt = (key & pmask) < prefix; eq = (key & pmask) == prefix;
            }
            const unsigned eqm = __ballot_sync(0xffffffffu, eq);
            const bool take_eq = eq && __popc(eqm & ((1u << lane) - 1u)) < ties_left;
            const unsigned tkm = __ballot_sync(0xffffffffu, take_eq);
            ties_left -= __popc(tkm);
            emit_chunk(K, L, __ballot_sync(0xffffffffu, lt) | tkm, e, i, j, d, lane);
        }
        if (L.cnt > 0) list_flush(K, L, e, i, lane, L.cnt);
        __syncwarp();
    }
}

static size_t scan_bytes(int64_t n) {
    size_t b = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b, (const int*)nullptr, (int64_t*)nullptr, (int)n);
    return b;
}

extern "C" int64_t cgvp_featurize_workspace_bytes(int64_t num_nodes, int64_t max_protein_len) {
    if (num_nodes < 0 || num_nodes >= ((int64_t)1 << 31) - 1) return -1;
    const int64_t ml = max_protein_len > 0 ? max_protein_len : 1;
    return align_up((num_nodes + 1) * 4, 256) + (int64_t)align_up((int64_t)scan_bytes(num_nodes + 1), 256) +
           align_up((2 * ml - 1) * 16 * 4, 256) + 256;
}

static int feat_checks(const float* ca, const int64_t* ptr, int64_t B, int64_t N, int32_t type) {
    CGVP_REQUIRE(B >= 0 && N >= 0 && N < ((int64_t)1 << 31) - 1, "featurize: bad sizes");
    CGVP_REQUIRE(N == 0 || (ca && ptr && B > 0), "featurize: null input");
    CGVP_REQUIRE(type >= 0 && type <= 2, "featurize: thresh_type must be 0 (dist), 1 (num) or 2 (prop)");
    return 0;
}

extern "C" int32_t cgvp_featurize_count(const float* ca, const int64_t* ptr, int64_t num_proteins, int64_t num_nodes,
                                        int64_t max_protein_len, double thresh, int32_t thresh_type,
                                        int32_t keep_self_loops, int64_t* row_offsets, void* ws, int64_t ws_bytes,
                                        cgvp_stream_t stream) {
    if (feat_checks(ca, ptr, num_proteins, num_nodes, thresh_type)) return -1;
    CGVP_REQUIRE(row_offsets, "featurize_count: null row_offsets");
    const int64_t need = cgvp_featurize_workspace_bytes(num_nodes, max_protein_len);
    CGVP_REQUIRE(ws && ws_bytes >= need, "featurize_count: workspace too small (%lld < %lld)", (long long)ws_bytes,
                 (long long)need);
    cudaStream_t st = (cudaStream_t)stream;
    FeatK K;
    memset(&K, 0, sizeof(K));
    K.ca = ca; K.ptr = ptr; K.B = num_proteins; K.N = num_nodes; K.thresh = thresh; K.type = thresh_type;
    K.keep_self = keep_self_loops != 0;
    K.deg = reinterpret_cast<int*>(ws);
    const int sms = cgvp_num_sms() > 0 ? cgvp_num_sms() : 148;
    int64_t blocks = cdiv64(num_nodes > 0 ? num_nodes : 1, FEAT_WARPS);
    if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
    feat_count_kernel<<<(int)blocks, FEAT_THREADS, 0, st>>>(K);
    CGVP_LAUNCH_CHECK("feat_count_kernel");
    void* tmp = reinterpret_cast<char*>(ws) + align_up((num_nodes + 1) * 4, 256);
    size_t tb = scan_bytes(num_nodes + 1);
    CGVP_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, (const int*)K.deg, row_offsets, (int)(num_nodes + 1), st));
    return 0;
}

extern "C" int32_t cgvp_featurize_fill(const float* ca, const int64_t* ptr, int64_t num_proteins, int64_t num_nodes,
                                       int64_t max_protein_len, double thresh, int32_t thresh_type,
                                       int32_t keep_self_loops, const int64_t* row_offsets, int64_t* edge_index,
                                       int64_t num_edges, float* edge_s, float* edge_v, void* ws, int64_t ws_bytes,
                                       cgvp_stream_t stream) {
    if (feat_checks(ca, ptr, num_proteins, num_nodes, thresh_type)) return -1;
    CGVP_REQUIRE(num_edges >= 0, "featurize_fill: bad edge count");
    if (num_edges == 0 || num_nodes == 0) return 0;
    CGVP_REQUIRE(row_offsets && edge_index && edge_s && edge_v, "featurize_fill: null output");
    FeatK K;
    memset(&K, 0, sizeof(K));
    K.ca = ca; K.ptr = ptr; K.B = num_proteins; K.N = num_nodes; K.thresh = thresh; K.type = thresh_type;
    K.keep_self = keep_self_loops != 0;
    K.row_offsets = row_offsets; K.ei = edge_index; K.E = num_edges; K.es = edge_s; K.ev = edge_v;
    const int sms = cgvp_num_sms() > 0 ? cgvp_num_sms() : 148;
    int64_t blocks = cdiv64(num_nodes, FEAT_WARPS);
    if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t ml = max_protein_len > 0 ? max_protein_len : 1;
    const int64_t need = cgvp_featurize_workspace_bytes(num_nodes, max_protein_len);
    if (ws && ws_bytes >= need) {          // positional-encoding table behind the count phase's scratch
        float* pe = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + align_up((num_nodes + 1) * 4, 256) +
                                             (int64_t)align_up((int64_t)scan_bytes(num_nodes + 1), 256));
        feat_pe_kernel<<<(int)cdiv64((2 * ml - 1) * 16, 256), 256, 0, st>>>(ml, pe);
        CGVP_LAUNCH_CHECK("feat_pe_kernel");
        K.pe = pe; K.max_len = ml;
    }
    size_t smem = 0;
    if (thresh_type != 0 && (size_t)FEAT_WARPS * ml * 8 <= 160 * 1024) {   // kNN: cache one row of fp64 distances per warp
        K.cache_rows = (int)ml;
        smem = (size_t)FEAT_WARPS * ml * 8;
        CGVP_CUDA(cudaFuncSetAttribute(feat_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int64_t per_sm = (int64_t)(200 * 1024) / (int64_t)(smem + 5 * 1024);
        if (blocks > (int64_t)sms * (per_sm > 0 ? per_sm : 1) * 4) blocks = (int64_t)sms * (per_sm > 0 ? per_sm : 1) * 4;
    }
    cgvp_prof_begin(CGVP_K_FEATURIZE, (cudaStream_t)stream);
    feat_fill_kernel<<<(int)blocks, FEAT_THREADS, smem, (cudaStream_t)stream>>>(K);
    cgvp_prof_end(CGVP_K_FEATURIZE, (cudaStream_t)stream);
    CGVP_LAUNCH_CHECK("feat_fill_kernel");
    return 0;
}

// ---- residue NODE features --------------------------------------------------------------------------------------------------
// compute_residue_node_features(vectorize_features=True, add_esm2_embeds=False), utils/create_protein_features.py:12-198.
// One thread per residue.  The reference works in numpy fp32 with separately rounded operations (no FMA contraction):
// differences, np.linalg.norm = sqrt((x*x + y*y) + z*z), np.cross = a*b - c*d, divide; the side-chain combination is
// evaluated in fp64 (np.sqrt(1/3) is a float64 scalar, :92) and cast to fp32 with the rest at create_graphs.py:21-23.
struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3_sub(V3 a, V3 b) { return {__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)}; }
__device__ __forceinline__ V3 v3_add(V3 a, V3 b) { return {__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z)}; }
__device__ __forceinline__ float v3_dot(V3 a, V3 b) {
    return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}
__device__ __forceinline__ V3 v3_cross(V3 a, V3 b) {
    return {__fsub_rn(__fmul_rn(a.y, b.z), __fmul_rn(a.z, b.y)), __fsub_rn(__fmul_rn(a.z, b.x), __fmul_rn(a.x, b.z)),
            __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x))};
}
__device__ __forceinline__ V3 v3_unit(V3 a) {                          // normalize_vecs, :360-365 (0 where the norm is 0)
    const float n = __fsqrt_rn(v3_dot(a, a));
    if (n == 0.f) return {0.f, 0.f, 0.f};
    return {__fdiv_rn(a.x, n), __fdiv_rn(a.y, n), __fdiv_rn(a.z, n)};
}
__device__ __forceinline__ V3 atom(const float* __restrict__ c, int64_t res, int a) {
    const float* p = c + (res * 4 + a) * 3;
    return {p[0], p[1], p[2]};
}

__global__ void node_feat_kernel(const float* __restrict__ coords, const int64_t* __restrict__ ptr, int64_t B, int64_t N,
                                 const int64_t* __restrict__ idents, const float* __restrict__ table, int num_types, int num_props,
                                 int posenc, float* __restrict__ out_s, float* __restrict__ out_v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int64_t b = find_protein(ptr, B, i);
    const int64_t lo = ptr[b], n = ptr[b + 1] - lo, li = i - lo;
    const int width = 6 + num_props + (posenc ? 16 : 0);
    float* s = out_s + i * width;
    // dihedrals (:34-66): backbone atoms N, CA, C of all residues in one chain; unit bond vectors u[k] = atom[k+1] - atom[k];
    // angle A[m] from (u[m], u[m+1], u[m+2]); residue li holds [0 ; A ; 0 ; 0][3 li .. 3 li + 2]
    float ang[3];
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        const int64_t e = 3 * li + t;
        ang[t] = 0.f;
        if (e >= 1 && e < 3 * n - 2) {
            const int64_t m = e - 1;                                     // needs atoms m .. m + 3
            V3 p[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) p[q] = atom(coords, lo + (m + q) / 3, (int)((m + q) % 3));
            const V3 u2 = v3_unit(v3_sub(p[1], p[0])), u1 = v3_unit(v3_sub(p[2], p[1])), u0 = v3_unit(v3_sub(p[3], p[2]));
            const V3 n1 = v3_unit(v3_cross(u1, u0)), n2 = v3_unit(v3_cross(u2, u1));
            const float c = fminf(fmaxf(v3_dot(n1, n2), -1.f), 1.f);
            const float sg = v3_dot(n1, u2);
            ang[t] = acosf(c) * (sg > 0.f ? 1.f : (sg < 0.f ? -1.f : 0.f));
        }
    }
#pragma unroll
    for (int t = 0; t < 3; ++t) { s[t] = cosf(ang[t]); s[3 + t] = sinf(ang[t]); }
    // amino-acid property columns (:95-109): a table look-up
    if (num_props > 0) {
        const int64_t id = idents[i];
        for (int j = 0; j < num_props; ++j) s[6 + j] = (id >= 0 && id < num_types) ? table[id * num_props + j] : 0.f;
    }
    // positional encoding of the residue index (:121-124, :368-385), fp64 like the reference
    if (posenc) {
        for (int k = 0; k < 8; ++k) {
            const double f = exp(2.0 * (double)k * -(log(10000.0) / 8.0));
            const double a = (double)li * f;
            s[6 + num_props + k] = (float)cos(a);
            s[6 + num_props + 8 + k] = (float)sin(a);
        }
    }
    // orientations (:69-78) and the virtual side chain (:81-92)
    const V3 ca = atom(coords, i, 1);
    V3 f = {0.f, 0.f, 0.f}, bk = {0.f, 0.f, 0.f};
    if (li + 1 < n) f = v3_unit(v3_sub(atom(coords, i + 1, 1), ca));
    if (li > 0) { const V3 t = v3_unit(v3_sub(ca, atom(coords, i - 1, 1))); bk = {-t.x, -t.y, -t.z}; }
    const V3 nv = v3_unit(v3_sub(atom(coords, i, 0), ca)), cv = v3_unit(v3_sub(atom(coords, i, 2), ca));
    const V3 bis = v3_unit(v3_add(nv, cv)), perp = v3_unit(v3_cross(cv, nv));
    const double k1 = sqrt(1.0 / 3.0), k2 = sqrt(2.0 / 3.0);
    float* v = out_v + i * 9;
    v[0] = f.x; v[1] = f.y; v[2] = f.z;
    v[3] = bk.x; v[4] = bk.y; v[5] = bk.z;
    v[6] = (float)__dsub_rn(__dmul_rn(-(double)bis.x, k1), __dmul_rn((double)perp.x, k2));
    v[7] = (float)__dsub_rn(__dmul_rn(-(double)bis.y, k1), __dmul_rn((double)perp.y, k2));
    v[8] = (float)__dsub_rn(__dmul_rn(-(double)bis.z, k1), __dmul_rn((double)perp.z, k2));
}

extern "C" int32_t cgvp_node_features(const float* res_coords, const int64_t* ptr, int64_t num_proteins, int64_t num_nodes,
                                      const int64_t* idents, const float* aa_table, int32_t num_types, int32_t num_props,
                                      int32_t add_posenc, float* out_s, float* out_v, void* stream) {
    CGVP_REQUIRE(num_proteins >= 0 && num_nodes >= 0, "node_features: bad sizes");
    CGVP_REQUIRE(num_props >= 0 && num_types >= 0, "node_features: bad table shape");
    if (num_nodes == 0) return 0;
    CGVP_REQUIRE(res_coords && ptr && out_s && out_v && num_proteins > 0, "node_features: null input");
    CGVP_REQUIRE(num_props == 0 || (idents && aa_table && num_types > 0), "node_features: property table without identities");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    node_feat_kernel<<<(unsigned)cdiv64(num_nodes, 128), 128, 0, st>>>(res_coords, ptr, num_proteins, num_nodes, idents, aa_table,
                                                                       num_types, num_props, add_posenc != 0, out_s, out_v);
    CGVP_LAUNCH_CHECK("node_feat_kernel");
    return 0;
}

// Library plumbing (error text, device queries) and the weight pack / gradient unpack kernels.
#include <stdarg.h>

#include "cgvp_common.cuh"

static thread_local char g_err[512] = "";

void cgvp_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cgvp_max_smem_optin() {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return 0;
    return v;
}

int cgvp_num_sms() {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    return v;
}

int cgvp_validate_gvp(const CgvpGvpDesc& d, const char* what) {
    CGVP_REQUIRE(d.si >= 0 && d.vi >= 0 && d.so > 0 && d.vo >= 0, "%s: bad GVP dims (%d,%d)->(%d,%d)", what, d.si, d.vi,
                 d.so, d.vo);
    CGVP_REQUIRE(d.si + d.vi > 0, "%s: GVP without inputs", what);
    CGVP_REQUIRE(d.vi == 0 || d.h > 0, "%s: h must be > 0 when vi > 0", what);
    CGVP_REQUIRE(d.scalar_act >= 0 && d.scalar_act <= 2 && d.vector_act >= 0 && d.vector_act <= 2,
                 "%s: unsupported activation code", what);
    return 0;
}

// ---- kernel timing -------------------------------------------------------------------------------------------------
#include <mutex>
#include <vector>
struct ProfRec { int id; cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
static bool g_prof_on = false;
static thread_local cudaEvent_t g_prof_open[CGVP_K_COUNT];

void cgvp_prof_begin(int id, cudaStream_t st) {
    if (!g_prof_on) return;
    cudaEvent_t a;
    if (cudaEventCreate(&a) != cudaSuccess) return;
    cudaEventRecord(a, st);
    g_prof_open[id] = a;
}
void cgvp_prof_end(int id, cudaStream_t st) {
    if (!g_prof_on || !g_prof_open[id]) return;
    cudaEvent_t b;
    if (cudaEventCreate(&b) != cudaSuccess) return;
    cudaEventRecord(b, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (g_prof.size() < 65536) g_prof.push_back({id, g_prof_open[id], b});
    g_prof_open[id] = nullptr;
}
extern "C" int32_t cgvp_profile_enable(int32_t on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
    g_prof_on = on != 0;
    return 0;
}
extern "C" int32_t cgvp_profile_collect(int32_t kernel_id, double* total_ms, int64_t* launches) {
    CGVP_REQUIRE(kernel_id >= 0 && kernel_id < CGVP_K_COUNT && total_ms && launches, "profile_collect: bad argument");
    std::lock_guard<std::mutex> lk(g_prof_mu);
    double tot = 0.0;
    int64_t n = 0;
    std::vector<ProfRec> keep;
    for (auto& r : g_prof) {
        if (r.id != kernel_id) { keep.push_back(r); continue; }
        float ms = 0.f;
        cudaEventSynchronize(r.b);
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { tot += ms; ++n; }
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    g_prof.swap(keep);
    *total_ms = tot; *launches = n;
    return 0;
}

extern "C" const char* cgvp_last_error(void) { return g_err; }
extern "C" int32_t cgvp_version(void) { return 100; }
extern "C" int32_t cgvp_sm_count(void) { return cgvp_num_sms(); }

extern "C" int64_t cgvp_gvp_packed_floats(const CgvpGvpDesc* desc) {
    if (!desc) return -1;
    return make_gvp_p(*desc).total_floats;
}

// ---------------------------------------------------------------------------------------------------------------
struct PackJob {
    GvpP g;
    CgvpGvpWeights w;
    float* dst;
};
#define PACK_MAX_JOBS 16
struct PackJobs {
    int n;
    PackJob j[PACK_MAX_JOBS];
};

__global__ void __launch_bounds__(256) pack_kernel(const PackJobs jobs) {
    const PackJob& J = jobs.j[blockIdx.x];
    const GvpP& g = J.g;
    const int hp = g.h4 * 4, sop = g.so4 * 4, vop = g.vo4 * 4, vip = g.vi4 * 4, ksd = g.si + g.h, ksdp = g.ksd4 * 4;
    for (int idx = threadIdx.x; idx < g.total_floats; idx += blockDim.x) {
        float v = 0.f;
        if (idx < g.o_ws_t) {                      // wh_t [vi_p][h_p]
            const int k = idx / hp, o = idx % hp;
            if (k < g.vi && o < g.h) v = J.w.wh[o * g.vi + k];
        } else if (idx < g.o_wv_t) {               // ws_t [ks_p][so_p]
            const int i = idx - g.o_ws_t, k = i / sop, o = i % sop;
            if (o < g.so) {
                if (k < ksd) v = J.w.ws[o * ksd + k];
                else if (k == ksd) v = J.w.bs[o];
            }
        } else if (idx < g.o_wsv_t) {              // wv_t [h_p][vo_p]
            const int i = idx - g.o_wv_t, k = i / vop, o = i % vop;
            if (k < g.h && o < g.vo) v = J.w.wv[o * g.h + k];
        } else if (idx < g.fwd_floats) {           // wsv_t [ksv_p][vo_p]
            const int i = idx - g.o_wsv_t, k = i / vop, o = i % vop;
            if (o < g.vo) {
                if (k < g.so) v = J.w.wsv[o * g.so + k];
                else if (k == g.so) v = J.w.bg[o];
            }
        } else if (idx < g.o_ws_b) {               // wh_b [h_p][vi_p]
            const int i = idx - g.o_wh_b, o = i / vip, k = i % vip;
            if (o < g.h && k < g.vi) v = J.w.wh[o * g.vi + k];
        } else if (idx < g.o_wv_b) {               // ws_b [so_p][ksd_p]
            const int i = idx - g.o_ws_b, o = i / ksdp, k = i % ksdp;
            if (o < g.so && k < ksd) v = J.w.ws[o * ksd + k];
        } else if (idx < g.o_wsv_b) {              // wv_b [vo_p][h_p]
            const int i = idx - g.o_wv_b, o = i / hp, k = i % hp;
            if (o < g.vo && k < g.h) v = J.w.wv[o * g.h + k];
        } else {                                   // wsv_b [vo_p][so_p]
            const int i = idx - g.o_wsv_b, o = i / sop, k = i % sop;
            if (o < g.vo && k < g.so) v = J.w.wsv[o * g.so + k];
        }
        J.dst[idx] = v;
    }
}

extern "C" int32_t cgvp_pack_weights(int32_t n, const CgvpGvpDesc* h_descs, const CgvpGvpWeights* h_weights,
                                     float* const* h_packed, cgvp_stream_t stream) {
    CGVP_REQUIRE(n >= 0 && h_descs && h_weights && h_packed, "cgvp_pack_weights: null argument");
    for (int base = 0; base < n; base += PACK_MAX_JOBS) {
        PackJobs jobs;
        jobs.n = n - base < PACK_MAX_JOBS ? n - base : PACK_MAX_JOBS;
        for (int i = 0; i < jobs.n; ++i) {
            const CgvpGvpDesc& d = h_descs[base + i];
            if (cgvp_validate_gvp(d, "cgvp_pack_weights")) return -1;
            GvpP g = make_gvp_p(d);
            const CgvpGvpWeights& w = h_weights[base + i];
            CGVP_REQUIRE(w.ws && w.bs, "cgvp_pack_weights: GVP %d has no ws/bias", base + i);
            CGVP_REQUIRE(g.vi == 0 || w.wh, "cgvp_pack_weights: GVP %d needs wh", base + i);
            CGVP_REQUIRE(!(g.vi > 0 && g.vo > 0) || w.wv, "cgvp_pack_weights: GVP %d needs wv", base + i);
            CGVP_REQUIRE(!g.gate || (w.wsv && w.bg), "cgvp_pack_weights: GVP %d needs wsv/bias", base + i);
            CGVP_REQUIRE(h_packed[base + i] && ((uintptr_t)h_packed[base + i] & 15) == 0,
                         "cgvp_pack_weights: packed block %d must be 16-byte aligned", base + i);
            jobs.j[i].g = g;
            jobs.j[i].w = w;
            jobs.j[i].dst = h_packed[base + i];
        }
        pack_kernel<<<jobs.n, 256, 0, (cudaStream_t)stream>>>(jobs);
        CGVP_LAUNCH_CHECK("pack_kernel");
    }
    return 0;
}

struct UnpackJob {
    GvpP g;
    CgvpGvpGrads w;
    const float* src;
};
struct UnpackJobs {
    int n;
    UnpackJob j[PACK_MAX_JOBS];
};

__global__ void __launch_bounds__(256) unpack_kernel(const UnpackJobs jobs) {
    const UnpackJob& J = jobs.j[blockIdx.x];
    const GvpP& g = J.g;
    const float* G = J.src;
    const int hp = g.h4 * 4, sop = g.so4 * 4, vop = g.vo4 * 4, ksd = g.si + g.h;
    const int n_wh = g.h * g.vi, n_ws = g.so * ksd, n_wv = (g.vi > 0 ? g.vo * g.h : 0), n_wsv = g.gate ? g.vo * g.so : 0;
    const int total = n_wh + n_ws + g.so + n_wv + n_wsv + (g.gate ? g.vo : 0);
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        int i = idx;
        if (i < n_wh) { const int o = i / g.vi, k = i % g.vi; J.w.wh[i] = G[g.o_wh_t + k * hp + o]; continue; }
        i -= n_wh;
        if (i < n_ws) { const int o = i / ksd, k = i % ksd; J.w.ws[i] = G[g.o_ws_t + k * sop + o]; continue; }
        i -= n_ws;
        if (i < g.so) { J.w.bs[i] = G[g.o_ws_t + ksd * sop + i]; continue; }
        i -= g.so;
        if (i < n_wv) { const int o = i / g.h, k = i % g.h; J.w.wv[i] = G[g.o_wv_t + k * vop + o]; continue; }
        i -= n_wv;
        if (i < n_wsv) { const int o = i / g.so, k = i % g.so; J.w.wsv[i] = G[g.o_wsv_t + k * vop + o]; continue; }
        i -= n_wsv;
        J.w.bg[i] = G[g.o_wsv_t + g.so * vop + i];
    }
}

extern "C" int32_t cgvp_unpack_grads(int32_t n, const CgvpGvpDesc* h_descs, const float* const* h_packed_grads,
                                     const CgvpGvpGrads* h_grads, cgvp_stream_t stream) {
    CGVP_REQUIRE(n >= 0 && h_descs && h_packed_grads && h_grads, "cgvp_unpack_grads: null argument");
    for (int base = 0; base < n; base += PACK_MAX_JOBS) {
        UnpackJobs jobs;
        jobs.n = n - base < PACK_MAX_JOBS ? n - base : PACK_MAX_JOBS;
        for (int i = 0; i < jobs.n; ++i) {
            if (cgvp_validate_gvp(h_descs[base + i], "cgvp_unpack_grads")) return -1;
            GvpP g = make_gvp_p(h_descs[base + i]);
            const CgvpGvpGrads& w = h_grads[base + i];
            CGVP_REQUIRE(w.ws && w.bs, "cgvp_unpack_grads: GVP %d has no ws/bias gradient buffers", base + i);
            CGVP_REQUIRE(g.vi == 0 || w.wh, "cgvp_unpack_grads: GVP %d needs wh", base + i);
            CGVP_REQUIRE(!(g.vi > 0 && g.vo > 0) || w.wv, "cgvp_unpack_grads: GVP %d needs wv", base + i);
            CGVP_REQUIRE(!g.gate || (w.wsv && w.bg), "cgvp_unpack_grads: GVP %d needs wsv/bias", base + i);
            jobs.j[i].g = g;
            jobs.j[i].w = w;
            jobs.j[i].src = h_packed_grads[base + i];
        }
        unpack_kernel<<<jobs.n, 256, 0, (cudaStream_t)stream>>>(jobs);
        CGVP_LAUNCH_CHECK("unpack_kernel");
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// out[i] = sum over the CTA partials.  Block = 8 warps x 32 outputs: warp w sums partials w, w + 8, ... of its lane's output
// (coalesced, several loads in flight); the eight slices are combined in warp order -- a fixed association, deterministic.
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partial, int nparts, int stride, int n, float* out) {
    __shared__ float sm[8][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;
    float s = 0.f;
    if (i < n) {
#pragma unroll 6
        for (int c = w; c < nparts; c += 8) s += __ldg(partial + (long long)c * stride + i);
    }
    sm[w][lane] = s;
    __syncthreads();
    if (w == 0 && i < n) {
        float t = 0.f;
#pragma unroll
        for (int ww = 0; ww < 8; ++ww) t += sm[ww][lane];
        out[i] = t;
    }
}

struct SegPack { CgvpSeg s[CGVP_MAX_SEGS]; };
__global__ void scatter_segments_kernel(const float* __restrict__ src, const SegPack segs) {
    const CgvpSeg& s = segs.s[blockIdx.y];
    if (!s.dst) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < s.n; i += gridDim.x * blockDim.x) s.dst[i] = src[s.off + i];
}

int cgvp_reduce_partials(const float* partial, int nparts, int stride, float* reduced, const CgvpSeg* segs, int nsegs,
                         cudaStream_t stream) {
    if (stride <= 0) return 0;
    CGVP_REQUIRE(nsegs <= CGVP_MAX_SEGS, "too many gradient segments");
    reduce_partials_kernel<<<cdiv(stride, 32), 256, 0, stream>>>(partial, nparts, stride, stride, reduced);
    CGVP_LAUNCH_CHECK("reduce_partials_kernel");
    if (nsegs > 0) {
        SegPack sp;
        memset(&sp, 0, sizeof(sp));
        for (int i = 0; i < nsegs; ++i) sp.s[i] = segs[i];
        scatter_segments_kernel<<<dim3(8, nsegs), 256, 0, stream>>>(reduced, sp);
        CGVP_LAUNCH_CHECK("scatter_segments_kernel");
    }
    return 0;
}

// Microbenchmark: where should the register-resident GVP kernels read their weights from?
//   smem  : shared memory, warp-uniform LDS.128 broadcast (what conv_reg.cu / rows_reg.cu did in round 1)
//   const : __constant__ bank, LDCU.128 into uniform registers, FFMA2 with a UR operand (no shared-memory crossbar
//           traffic, no vector registers for weights)
// Work: the full forward of message GVP 0 at checkpoint dims (64,9)->(16,4), h=9, per row (1 583 FMA), ITER rows per thread.
// Build (from the repo root; header-only dependency on cgvp_reg.cuh):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I caster_dta_b200/csrc \
//        scripts/microbench/weights_src.cu -o scripts/microbench/weights_src
// Result on a B200: profiles/r2_weights_src_microbench.jsonl
#include <cstdio>
#include <vector>
#include "cgvp_reg.cuh"
void cgvp_set_error(const char*, ...) {}
using namespace cgvpr;
using G0 = GvpC<64, 9, 16, 4, 9, CGVP_ACT_RELU, CGVP_ACT_NONE, 1>;
__constant__ float c_w[G0::TOTAL_FLOATS];

template <bool CONST, int MINB>
__global__ void __launch_bounds__(256, MINB) k(const float* __restrict__ wg, float* __restrict__ out, int iters) {
    extern __shared__ __align__(16) float sm[];
    const float* W;
    if constexpr (CONST) {
        W = c_w;
    } else {
        for (int i = threadIdx.x; i < G0::FWD_FLOATS; i += blockDim.x) sm[i] = wg[i];
        __syncthreads();
        W = sm;
    }
    float acc = 0.f;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        asm volatile("" ::: "memory");          // the weight loads must stay inside the loop (as in the tile loop of the kernels)
        float s0[1][G0::SI], v0[3][G0::VI1], s1[1][G0::SO], v1[3][G0::VO1];
#pragma unroll
        for (int i = 0; i < G0::SI; ++i) s0[0][i] = __int_as_float(0x3f000000 + ((tid * 31 + it * 7 + i * 13) & 0xffff));
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int i = 0; i < G0::VI; ++i) v0[p][i] = __int_as_float(0x3f000000 + ((tid * 17 + it * 5 + i * 3 + p) & 0xffff));
        Save<G0> sv;
        gvp_fwd<G0>(W, s0, v0, s1, v1, sv);
#pragma unroll
        for (int i = 0; i < G0::SO; ++i) acc += s1[0][i];
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int i = 0; i < G0::VO; ++i) acc += v1[p][i];
    }
    out[tid] = acc;
}

template <bool CONST, int MINB>
static void run(const char* name, const float* wg, float* out, int grid, int iters) {
    size_t smem = CONST ? 0 : G0::FWD_FLOATS * 4;
    cudaFuncSetAttribute(k<CONST, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<CONST, MINB><<<grid, 256, smem>>>(wg, out, 2);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<CONST, MINB><<<grid, 256, smem>>>(wg, out, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double rows = (double)grid * 256 * iters;
    const double fma = rows * 1583.0;
    cudaFuncAttributes at;
    cudaFuncGetAttributes(&at, k<CONST, MINB>);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<CONST, MINB>, 256, smem);
    printf("{\"variant\": \"%s\", \"min_blocks\": %d, \"regs\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"grows_per_s\": %.3f, \"tflops\": %.2f, \"err\": \"%s\"}\n",
           name, MINB, at.numRegs, occ, ms, rows / ms / 1e6, 2 * fma / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    std::vector<float> w(G0::TOTAL_FLOATS);
    for (size_t i = 0; i < w.size(); ++i) w[i] = 0.01f * (float)((i * 7919) % 101 - 50);
    float *wg, *out;
    cudaMalloc(&wg, w.size() * 4);
    cudaMemcpy(wg, w.data(), w.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(c_w, w.data(), w.size() * 4);
    const int grid = 148 * 8, iters = 64;
    cudaMalloc(&out, (size_t)grid * 256 * 4);
    run<false, 1>("smem", wg, out, grid, iters);
    run<false, 2>("smem", wg, out, grid, iters);
    run<false, 3>("smem", wg, out, grid, iters);
    run<true, 1>("const", wg, out, grid, iters);
    run<true, 2>("const", wg, out, grid, iters);
    run<true, 3>("const", wg, out, grid, iters);
    return 0;
}

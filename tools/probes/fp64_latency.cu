// FP64 dependent-issue latency and per-scheduler throughput on the device (probe, not product):
//   nvcc -O3 -fmad=false -gencode arch=compute_100a,code=sm_100a -o build/fp64_latency tools/probes/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS, int OP>
__global__ void k(double a, double b, int iters, double* out, long long* cyc) {
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = a + c + threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) {
                if (OP == 0) x[c] = __fma_rn(x[c], b, a);
                else if (OP == 1) x[c] = __dadd_rn(x[c], b);
                else x[c] = __dmul_rn(x[c], b);
            }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int CHAINS, int OP>
void run(const char* name, int threads) {
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    k<CHAINS, OP><<<1, threads>>>(1.0000001, 0.9999999, iters, out, cyc);
    k<CHAINS, OP><<<1, threads>>>(1.0000001, 0.9999999, iters, out, cyc);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)h / (iters * 8.0);
    printf("%s threads %4d (warps/scheduler %d) chains %d: %.2f cycles per round of %d ops per warp -> %.2f cycles/op/warp, scheduler issues one op per %.2f cycles\n",
           name, threads, threads / 128 > 0 ? threads / 128 : 1, CHAINS, per, CHAINS, per / CHAINS, per / CHAINS / (threads >= 128 ? threads / 128 : 1));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<1, 0>("DFMA", 32); run<1, 1>("DADD", 32); run<1, 2>("DMUL", 32);
    run<2, 0>("DFMA", 32); run<4, 0>("DFMA", 32); run<8, 0>("DFMA", 32);
    run<1, 0>("DFMA", 128); run<1, 0>("DFMA", 256); run<1, 0>("DFMA", 512); run<1, 0>("DFMA", 1024);
    run<2, 0>("DFMA", 512); run<4, 0>("DFMA", 512);
    return 0;
}

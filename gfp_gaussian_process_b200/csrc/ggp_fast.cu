// ggp_fast.cu — the FAST likelihood kernels (ggp_fast.cuh): one thread per (cell, parameter vector), one launch per
// generation, mother -> daughter hand-over through the same SoA state buffer and the same fixed-order reduction as the strict
// path.  Compiled with FMA contraction ON (-fmad=true); nothing of the strict math is included here.
//
// Build: nvcc -std=c++17 -O3 -fmad=true -gencode arch=compute_100a,code=sm_100a -lineinfo -c
#include "ggp_fast_api.h"
#include "ggp_fast.cuh"

#define GGP_FAST_BLOCK 128

template <int N>
__global__ void __launch_bounds__(GGP_FAST_BLOCK) ggp_fast_loglik_kernel(const GgpDevForest F, const GgpFwdArgs A, int* __restrict__ invalid) {
    __shared__ double sp[GGP_NP];
    __shared__ double red[GGP_FAST_BLOCK / 32];
    const int lane_slot = blockIdx.x * GGP_FAST_BLOCK + threadIdx.x;
    const bool active = lane_slot < A.n_slots;
    const int slot = A.slot0 + (active ? lane_slot : 0);
    const int v = blockIdx.y;
    if (threadIdx.x < GGP_NP)
        sp[threadIdx.x] = A.params ? A.params[(int64_t)(A.v0 + v) * GGP_NP + threadIdx.x] : A.inline_params[(A.v0 + v) * GGP_NP + threadIdx.x];
    __syncthreads();
    double own = 0.0;
    if (active) {
        const int64_t vstride = (int64_t)A.v_count * F.n_cells;
        const int64_t vbase = (int64_t)v * F.n_cells;
        const int parent = F.s_parent[slot];
        GgpFastState<double> s;
        if (parent >= 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) s.m[k] = A.state[k * vstride + vbase + parent];
#pragma unroll
            for (int k = 0; k < 10; ++k) s.c[k] = A.state[(4 + k) * vstride + vbase + parent];
        }
        GgpFastConsts<double, N> K;
        K.t = __longlong_as_double(0x7ff8000000000000ll);
        bool valid = true;
        own = ggp_fast_cell<double, N>(F, slot, sp, s, K, GgpGLRule<N>(), valid);
        if (F.s_d1[slot] >= 0 || F.s_d2[slot] >= 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) A.state[k * vstride + vbase + slot] = s.m[k];
#pragma unroll
            for (int k = 0; k < 10; ++k) A.state[(4 + k) * vstride + vbase + slot] = s.c[k];
        }
        if (A.cell_ll) A.cell_ll[(int64_t)(A.v0 + v) * F.n_cells + F.s_cell[slot]] = own;
        if (!valid) invalid[A.v0 + v] = 1;
    }
    // fixed-order reduction: xor-shuffle tree inside the warp, then the warps in index order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) own = own + __shfl_xor_sync(0xffffffffu, own, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = own;
    __syncthreads();
    if (threadIdx.x == 0) {
        double bs = 0.0;
        for (int i = 0; i < GGP_FAST_BLOCK / 32; ++i) bs = bs + red[i];
        A.partial[(int64_t)v * A.n_partial + A.partial0 + blockIdx.x] = bs;
    }
}

bool ggp_fast_supported_nodes(int n) { return n == 4 || n == 5 || n == 6 || n == 8 || n == 10; }

cudaError_t ggp_fast_loglik_launch(const GgpDevForest& F, const GgpFwdArgs& A, int* invalid, int n_nodes, cudaStream_t stream) {
    const dim3 grid((unsigned)((A.n_slots + GGP_FAST_BLOCK - 1) / GGP_FAST_BLOCK), (unsigned)A.v_count);
    switch (n_nodes) {
        case 4: ggp_fast_loglik_kernel<4><<<grid, GGP_FAST_BLOCK, 0, stream>>>(F, A, invalid); break;
        case 5: ggp_fast_loglik_kernel<5><<<grid, GGP_FAST_BLOCK, 0, stream>>>(F, A, invalid); break;
        case 6: ggp_fast_loglik_kernel<6><<<grid, GGP_FAST_BLOCK, 0, stream>>>(F, A, invalid); break;
        case 8: ggp_fast_loglik_kernel<8><<<grid, GGP_FAST_BLOCK, 0, stream>>>(F, A, invalid); break;
        case 10: ggp_fast_loglik_kernel<10><<<grid, GGP_FAST_BLOCK, 0, stream>>>(F, A, invalid); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

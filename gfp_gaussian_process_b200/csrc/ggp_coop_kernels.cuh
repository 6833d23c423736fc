// ggp_coop_kernels.cuh — lineage-forest passes with the four-warps-per-32-cells step of ggp_coop.cuh.
//
// Replaces, from the reference (paths under src/): likelihood_recr / total_likelihood (likelihood.h:110-174).
// Same launch structure as ggp_kernels.cuh (one launch per generation, y = parameter vector), but a block of
// 128 threads owns 32 cells: lane = cell, warp = role.  See ggp_coop.cuh for the phase plan.
#pragma once
#include "ggp_kernels.cuh"
#include "ggp_coop.cuh"

// A block holds NG groups of four warps (32 cells each) that share one copy of the math tables and advance phase by
// phase behind the same block barriers.  Warp w has role w % 4 (so the four warps of one scheduler run the same
// role's instruction stream at about the same time: one instruction fetch serves NG groups) and group w / 4.
//   NG = 1: 128 threads, 4 blocks per SM; finest granularity, used for generations with few cells
//   NG = 4: 512 threads, 1 block per SM; 4x fewer instruction-cache fills per cell (the step's hot code is ~75 kB,
//           over twice the 32 kB L1.5 instruction cache, so every step streams from L2)
#define GGP_COOP_BLOCK(NG) ((NG) * GGP_COOP_ROLES * 32)
#define GGP_COOP_SMEM_BYTES(NG) (sizeof(GgpMathTables) + (size_t)(NG) * GGP_CS_COUNT * GGP_COOP_CELLS * sizeof(double))

template <int NG>
__global__ void __launch_bounds__(GGP_COOP_BLOCK(NG), 4 / NG) ggp_loglik_coop_kernel(const GgpDevForest F, const GgpFwdArgs A) {
    GgpMathTables& T = *reinterpret_cast<GgpMathTables*>(ggp_smem);
    __shared__ double sp[GGP_NP];
    __shared__ int s_steps[NG * GGP_COOP_ROLES];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int role = warp & (GGP_COOP_ROLES - 1), group = warp / GGP_COOP_ROLES;
    const int v = blockIdx.y;
    if (threadIdx.x < GGP_NP) sp[threadIdx.x] = A.params[(int64_t)(A.v0 + v) * GGP_NP + threadIdx.x];
    ggp_stage_tables(&T);   // ends with a block barrier
    GgpScratch S;
    S.base = reinterpret_cast<double*>(ggp_smem + sizeof(GgpMathTables)) + (size_t)group * GGP_CS_COUNT * GGP_COOP_CELLS + lane;
    S.stride = GGP_COOP_CELLS;

    const int gidx = blockIdx.x * NG + group;                  // 32-cell group inside this generation
    const int in_gen = gidx * GGP_COOP_CELLS + lane;
    const bool active = in_gen < A.n_slots;
    const int slot = A.slot0 + (active ? in_gen : 0);
    int64_t off = 0;
    int n = 0, parent = -1;
    if (active) {
        off = F.s_off[slot];
        n = F.s_n[slot];
        parent = F.s_parent[slot];
    }
    const int64_t vstride = (int64_t)A.v_count * F.n_cells, vbase = (int64_t)v * F.n_cells;
    const double* p = sp;
    double own = 0.0;
    int t = 0;
    int64_t from = off;
    if (active) {
        if (parent < 0) {
            if (role == 0) {   // root: first update on the full matrix (predictions.h:63-82, likelihood.h:53-69)
                double mu[4], C[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) C[i] = 0.0;
                mu[0] = F.init_f[0]; mu[1] = F.init_f[1];
                C[0] = F.init_f[2];  C[5] = F.init_f[3];
                mu[2] = p[0]; mu[3] = p[3];
                C[10] = p[2] / (2. * p[1]);
                C[15] = p[5] / (2. * p[4]);
                const GgpMeas m = ggp_measure16(mu, C, F.x[off], F.g[off], p[7], p[8], F.model);
                const double ll = ggp_log_evidence(m, &T);
                own = own + ll;
                if (ll != ll) GGP_NAN_MIN(A.nan_key + A.v0 + v, F.s_dfs0[slot]);
                ggp_posterior16(mu, C, m);
                GgpState s;
                ggp_state_from16(s, mu, C);
#pragma unroll
                for (int k = 0; k < 4; ++k) S[GGP_CS_ST + k] = s.m[k];
#pragma unroll
                for (int k = 0; k < 10; ++k) S[GGP_CS_ST + 4 + k] = s.c[k];
            }
        } else {   // daughter: mother's last posterior; the division gap is step "-1"
            for (int k = role; k < 14; k += GGP_COOP_ROLES) S[GGP_CS_ST + k] = A.state[k * vstride + vbase + parent];
            t = -1;
            from = F.s_off[parent] + F.s_n[parent] - 1;
        }
    }
    const int steps = active ? n - 1 - t : 0;
    int max_steps = __reduce_max_sync(0xffffffffu, steps);
    if (NG > 1) {   // the barriers are block wide: every group runs the block's longest cell
        if (lane == 0) s_steps[warp] = max_steps;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < NG * GGP_COOP_ROLES; ++i) max_steps = max(max_steps, s_steps[i]);
    }
    const GgpOuParams ou = ggp_ou(p, false);
    // measurements of the next step are fetched one step ahead (the loads retire behind a whole step of arithmetic)
    double dt = 0.0, xo = 0.0, go = 0.0;
    if (steps > 0) {
        const int64_t at = off + t + 1;
        dt = F.time[at] - F.time[from];
        xo = F.x[at];
        go = F.g[at];
    }
    __syncthreads();
    for (int it = 0; it < max_steps; ++it) {
        const bool live = it < steps;
        double dt_n = 0.0, xo_n = 0.0, go_n = 0.0;
        if (it + 1 < steps) {
            const int64_t at = off + t + 2;
            dt_n = F.time[at] - F.time[at - 1];
            xo_n = F.x[at];
            go_n = F.g[at];
        }
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) {
            if (live) ggp_coop_run_phase(ph, role, S, ou, dt, &T);
            __syncthreads();
        }
        if (live) {
            const double ll = ggp_coop_ph4(role, S, t < 0, p, xo, go, F.model, &T);
            ++t;
            if (role == 0) {
                own = own + ll;
                if (ll != ll) GGP_NAN_MIN(A.nan_key + A.v0 + v, F.s_dfs0[slot] + t);
            }
        }
        dt = dt_n; xo = xo_n; go = go_n;
        __syncthreads();
    }
    if (active && (F.s_d1[slot] >= 0 || F.s_d2[slot] >= 0)) {
        for (int k = role; k < 14; k += GGP_COOP_ROLES) A.state[k * vstride + vbase + slot] = S[GGP_CS_ST + k];
    }
    if (role == 0) {
        if (active && A.cell_ll) A.cell_ll[(int64_t)(A.v0 + v) * F.n_cells + F.s_cell[slot]] = own;
        double bs = own;   // fixed-order reduction over the block's 32 cells
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bs = bs + __shfl_xor_sync(0xffffffffu, bs, o);
        if (lane == 0 && gidx * GGP_COOP_CELLS < A.n_slots) A.partial[(int64_t)v * A.n_partial + A.partial0 + gidx] = bs;
    }
}

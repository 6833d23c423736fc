// ggp_layout.hpp — host-side flattening of a lineage forest into generation-ordered slot arrays.
// Pure C++ (no CUDA) so the product library and the CPU host check share it.
//
// Host side of what the reference does in moma_input.h:125-151 (build_cell_genealogy), :177-189
// (get_roots), :663-735 (init_cells_f/r statistics), likelihood.h:110-122 (depth-first order, used here
// only to report the first NaN) and predictions.h:482-493 (segment used by combine_predictions).
#pragma once
#include <stdint.h>

#include <algorithm>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/ggp_b200.h"

struct GgpLayout {
    int64_t n_cells = 0, n_ctp = 0;
    int32_t n_roots = 0, n_gen = 0, max_seg = 0;
    std::vector<int64_t> gen_start;          // [n_gen+1] slot ranges
    // streamed upload: the series are cut into n_chunks contiguous ctp ranges; a tree belongs to the chunk its last
    // point arrives with, and inside a generation the slots are grouped by chunk, so a chunk's trees can be evaluated
    // (all generations) as soon as that chunk has landed on the device
    int32_t n_chunks = 1;
    std::vector<int64_t> ctp_chunk_start;    // [n_chunks+1]
    std::vector<int64_t> gen_chunk_start;    // [n_gen][n_chunks+1] slot ranges
    std::vector<int32_t> slot_of_cell, cell_of_slot;
    std::vector<int32_t> dfs_cells;          // cells in the reference's depth-first order
    std::vector<int64_t> dfs_ctp0;           // [n_cells+1] by dfs position
    // per slot
    std::vector<int64_t> s_off, s_dfs0;
    std::vector<int32_t> s_n, s_parent, s_d1, s_d2, s_root, s_cell;
    // per ctp
    std::vector<int32_t> seg, comb_seg, ctp_slot;
    double init_f[4] = {0, 0, 0, 0}, init_r[4] = {0, 0, 0, 0};

    // init_cells_f / init_cells_r (moma_input.h:663-735): left-to-right sums over the cells with more than
    // one point, var = E[x^2] - E[x]^2
    static void init_stats(const ggp_forest_desc* d, double* f4, double* r4) {
        for (int dir = 0; dir < 2; ++dir) {
            double sx = 0, sg = 0, sxx = 0, sgg = 0;
            int64_t cnt = 0;
            for (int64_t c = 0; c < d->n_cells; ++c) {
                const int64_t o = d->cell_offset[c], n = d->cell_offset[c + 1] - o;
                if (n > 1) {
                    const int64_t k = dir == 0 ? o : o + n - 1;
                    sx += d->log_length[k];
                    sg += d->fp[k];
                    sxx += d->log_length[k] * d->log_length[k];
                    sgg += d->fp[k] * d->fp[k];
                    ++cnt;
                }
            }
            double* out = dir == 0 ? f4 : r4;
            const double mx = sx / cnt, mg = sg / cnt;
            out[0] = mx;
            out[1] = mg;
            out[2] = sxx / cnt - mx * mx;
            out[3] = sgg / cnt - mg * mg;
        }
    }

    // returns "" on success, else the reason (the reference throws std::invalid_argument)
    // chunk_fractions (optional, want_chunks entries, any positive scale): relative sizes of the upload chunks
    std::string build(const ggp_forest_desc* d, int32_t want_chunks = 1, const double* chunk_fractions = nullptr) {
        if (!d) return "null descriptor";
        if (d->n_cells <= 0 || d->n_ctp <= 0 || !d->cell_offset || !d->parent || !d->time || !d->log_length || !d->fp)
            return "empty forest or missing array";
        if (d->n_cells > INT32_MAX) return "more than 2^31-1 cells in one handle";
        if (d->cell_offset[0] != 0 || d->cell_offset[d->n_cells] != d->n_ctp) return "cell_offset does not span [0, n_ctp]";
        const int64_t N = d->n_cells;
        n_cells = N;
        n_ctp = d->n_ctp;
        for (int64_t c = 0; c < N; ++c) {
            if (d->cell_offset[c + 1] <= d->cell_offset[c]) return "cell without time points";
            if (d->parent[c] >= N || d->parent[c] == c) return "parent index out of range";
        }
        // daughters: from the caller if given, else the first two children in cell order
        std::vector<int32_t> d1(N, -1), d2(N, -1);
        if (d->daughter1 && d->daughter2) {
            std::copy(d->daughter1, d->daughter1 + N, d1.begin());
            std::copy(d->daughter2, d->daughter2 + N, d2.begin());
            for (int64_t c = 0; c < N; ++c)
                for (int32_t k : {d1[c], d2[c]})
                    if (k >= N || (k >= 0 && d->parent[k] != c)) return "daughter link does not match parent link";
        } else {
            for (int64_t c = 0; c < N; ++c) {
                const int32_t p = d->parent[c];
                if (p < 0) continue;
                if (d1[p] < 0) d1[p] = (int32_t)c;
                else if (d2[p] < 0) d2[p] = (int32_t)c;
                else return "both daughter pointers are set";   // build_cell_genealogy's error, moma_input.h:141-147
            }
        }
        // generation of every cell (parents may come after daughters in cell order)
        std::vector<int32_t> gen(N, -1);
        {
            std::vector<int32_t> chain;
            for (int64_t c = 0; c < N; ++c) {
                if (gen[c] >= 0) continue;
                chain.clear();
                int64_t u = c;
                while (u >= 0 && gen[u] < 0) {
                    chain.push_back((int32_t)u);
                    if ((int64_t)chain.size() > N) return "cycle in parent links";
                    u = d->parent[u];
                }
                int32_t base = u < 0 ? -1 : gen[u];
                for (auto it = chain.rbegin(); it != chain.rend(); ++it) gen[*it] = ++base;
            }
        }
        // a cell that is nobody's daughter1/2 is unreachable in the reference's recursion; refuse rather than differ
        for (int64_t c = 0; c < N; ++c) {
            const int32_t p = d->parent[c];
            if (p >= 0 && d1[p] != c && d2[p] != c) return "cell has a parent but is not one of its two daughters";
        }
        n_gen = *std::max_element(gen.begin(), gen.end()) + 1;
        // upload chunks and the chunk of every cell (= of its tree)
        n_chunks = std::max<int32_t>(1, std::min<int64_t>(want_chunks, d->n_ctp));
        ctp_chunk_start.assign(n_chunks + 1, 0);
        if (chunk_fractions && n_chunks == want_chunks) {
            double tot = 0.0, acc = 0.0;
            for (int32_t k = 0; k < n_chunks; ++k) tot += chunk_fractions[k];
            for (int32_t k = 0; k < n_chunks; ++k) {
                ctp_chunk_start[k] = std::max<int64_t>(k ? ctp_chunk_start[k - 1] : 0, (int64_t)((double)d->n_ctp * (acc / tot)));
                acc += chunk_fractions[k];
            }
            ctp_chunk_start[n_chunks] = d->n_ctp;
        } else {
            for (int32_t k = 0; k <= n_chunks; ++k) ctp_chunk_start[k] = d->n_ctp * k / n_chunks;
        }
        std::vector<int32_t> chunk(N, 0);
        if (n_chunks > 1) {
            std::vector<int32_t> root_of(N, -1), order(N);
            std::iota(order.begin(), order.end(), 0);
            std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return gen[a] < gen[b]; });
            std::vector<int32_t> tree_chunk(N, 0);
            for (int32_t c : order) {   // parents before daughters
                root_of[c] = d->parent[c] < 0 ? c : root_of[d->parent[c]];
                const int64_t last = d->cell_offset[c + 1] - 1;
                const int32_t k = (int32_t)(std::upper_bound(ctp_chunk_start.begin(), ctp_chunk_start.end(), last) - ctp_chunk_start.begin()) - 1;
                tree_chunk[root_of[c]] = std::max(tree_chunk[root_of[c]], k);
            }
            for (int64_t c = 0; c < N; ++c) chunk[c] = tree_chunk[root_of[c]];
        }
        // slots: by generation, then upload chunk, then number of points (descending), then cell index
        cell_of_slot.resize(N);
        std::iota(cell_of_slot.begin(), cell_of_slot.end(), 0);
        std::stable_sort(cell_of_slot.begin(), cell_of_slot.end(), [&](int32_t a, int32_t b) {
            if (gen[a] != gen[b]) return gen[a] < gen[b];
            if (chunk[a] != chunk[b]) return chunk[a] < chunk[b];
            const int64_t na = d->cell_offset[a + 1] - d->cell_offset[a], nb = d->cell_offset[b + 1] - d->cell_offset[b];
            if (na != nb) return na > nb;
            return a < b;
        });
        slot_of_cell.assign(N, 0);
        for (int64_t s = 0; s < N; ++s) slot_of_cell[cell_of_slot[s]] = (int32_t)s;
        gen_start.assign(n_gen + 1, 0);
        for (int64_t c = 0; c < N; ++c) gen_start[gen[c] + 1]++;
        for (int g = 0; g < n_gen; ++g) gen_start[g + 1] += gen_start[g];
        gen_chunk_start.assign((size_t)n_gen * (n_chunks + 1), 0);
        for (int64_t c = 0; c < N; ++c) gen_chunk_start[(size_t)gen[c] * (n_chunks + 1) + chunk[c] + 1]++;
        for (int g = 0; g < n_gen; ++g) {
            int64_t* row = gen_chunk_start.data() + (size_t)g * (n_chunks + 1);
            row[0] = gen_start[g];
            for (int k = 0; k < n_chunks; ++k) row[k + 1] += row[k];
        }
        // depth-first order of the reference (roots in cell order; cell, daughter1 subtree, daughter2 subtree)
        std::vector<int64_t> dfs0_of_cell(N, 0);
        {
            dfs_cells.clear();
            dfs_ctp0.clear();
            std::vector<int32_t> st;
            int64_t rank = 0;
            for (int64_t r = 0; r < N; ++r) {
                if (d->parent[r] >= 0) continue;
                st.push_back((int32_t)r);
                while (!st.empty()) {
                    const int32_t u = st.back();
                    st.pop_back();
                    dfs_cells.push_back(u);
                    dfs_ctp0.push_back(rank);
                    dfs0_of_cell[u] = rank;
                    rank += d->cell_offset[u + 1] - d->cell_offset[u];
                    if (d2[u] >= 0) st.push_back(d2[u]);
                    if (d1[u] >= 0) st.push_back(d1[u]);
                }
            }
            dfs_ctp0.push_back(rank);
            if ((int64_t)dfs_cells.size() != N) return "daughter links do not form a forest";
        }
        s_off.resize(N); s_dfs0.resize(N); s_n.resize(N); s_parent.resize(N);
        s_d1.resize(N); s_d2.resize(N); s_root.resize(N); s_cell.resize(N);
        std::vector<int32_t> root_no(N, -1);
        n_roots = 0;
        for (int64_t c = 0; c < N; ++c)
            if (d->parent[c] < 0) root_no[c] = n_roots++;
        for (int64_t s = 0; s < N; ++s) {
            const int32_t c = cell_of_slot[s];
            s_off[s] = d->cell_offset[c];
            s_n[s] = (int32_t)(d->cell_offset[c + 1] - d->cell_offset[c]);
            s_parent[s] = d->parent[c] < 0 ? -1 : slot_of_cell[d->parent[c]];
            s_d1[s] = d1[c] < 0 ? -1 : slot_of_cell[d1[c]];
            s_d2[s] = d2[c] < 0 ? -1 : slot_of_cell[d2[c]];
            s_root[s] = root_no[c];
            s_cell[s] = c;
            s_dfs0[s] = dfs0_of_cell[c];
        }
        // segments, and the segment combine_predictions divides by (predictions.h:482-493)
        seg.assign(d->n_ctp, 0);
        if (d->segment) std::copy(d->segment, d->segment + d->n_ctp, seg.begin());
        max_seg = 0;
        for (int64_t i = 0; i < d->n_ctp; ++i) {
            if (seg[i] < 0) return "negative segment index";
            max_seg = std::max(max_seg, seg[i]);
        }
        ctp_slot.assign(d->n_ctp, 0);
        for (int64_t s = 0; s < N; ++s)
            for (int64_t k = s_off[s]; k < s_off[s] + s_n[s]; ++k) ctp_slot[k] = (int32_t)s;
        comb_seg = seg;
        for (int64_t c = 0; c < N; ++c)
            if (d->parent[c] >= 0) comb_seg[d->cell_offset[c]] = seg[d->cell_offset[d->parent[c] + 1] - 1];
        if (d->compute_init) init_stats(d, init_f, init_r);
        else for (int i = 0; i < 4; ++i) { init_f[i] = d->init_f[i]; init_r[i] = d->init_r[i]; }
        return "";
    }

    // (cell, time index) of a depth-first ctp rank
    void locate(int64_t rank, int64_t* cell, int64_t* t) const {
        const auto it = std::upper_bound(dfs_ctp0.begin(), dfs_ctp0.end(), rank);
        const int64_t pos = (it - dfs_ctp0.begin()) - 1;
        *cell = dfs_cells[pos];
        *t = rank - dfs_ctp0[pos];
    }
};

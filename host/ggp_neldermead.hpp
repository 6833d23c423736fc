// host/ggp_neldermead.hpp — bound-constrained Nelder-Mead simplex minimiser.
//
// The reference maximises the likelihood with NLopt 2.7.1's LN_NELDERMEAD (minimizer_nlopt.h:59-87, 154-192:
// set_lower/upper_bounds, set_initial_step, set_ftol_abs; fixed parameters have lb == ub).  NLopt is not available
// in this environment (no source, no wheel), so its published algorithm is restated here from the NLopt
// documentation and J. A. Nelder & R. Mead (1965) with the Richardson & Kuester (1973) bound handling NLopt uses:
//   * dimensions with lb == ub are eliminated before the search;
//   * initial simplex x0, x0 + step_i e_i (a vertex that would leave the box is put on the bound if the bound is
//     further than 0.1 |step_i| away, else stepped in the other direction);
//   * reflection (1), expansion (2), contraction (1/2, inside if f_high <= f_reflected else outside),
//     shrink towards the best vertex (1/2); every new point is clipped ("pinned") to the box;
//   * stop when f_high - f_low < ftol_abs, or when a new point coincides with the centroid / the old point.
// Iterate-by-iterate parity with NLopt is UNPINNED (nothing to compare against here); optima are compared.
//
// What is new: the objective takes a BATCH of points.  The n+1 evaluations of the initial simplex and the n
// evaluations of a shrink are independent and go to the GPU as one launch; with `speculate` the reflection,
// expansion and both contractions of an iteration are evaluated together as well (they depend only on the
// simplex geometry), which turns the ~2 dependent launches per iteration into one.  With a larger `spec_batch`
// the batch also holds the candidates of the NEXT iterations for every way the current one can end (which point is
// accepted, which vertex is then the worst): on a small data set a launch costs the same for 1 or 200 parameter
// vectors (one lineage tree is a chain of dependent steps), so several iterations are served by one launch.  The
// points are looked up in a cache when the sequential algorithm asks for them; a miss costs one more launch, never
// a different result.  Evaluation ORDER as seen by the objective's log stays the sequential algorithm's.
#pragma once
#include <algorithm>
#include <cmath>
#include <deque>
#include <functional>
#include <map>
#include <set>
#include <vector>

namespace ggp {

struct NelderMeadResult {
    std::vector<double> x;
    double f = 0;
    int evaluations = 0;   // evaluations the sequential algorithm consumed
    int launches = 0;      // batched objective calls
    const char* reason = "";
};

// evaluate(points, record): values of the points, in order.  record = false: a speculative batch whose members are
// reported later, one by one, through commit(x, f) if the algorithm uses them.
struct BatchObjective {
    std::function<std::vector<double>(const std::vector<std::vector<double>>&, bool record)> evaluate;
    std::function<void(const std::vector<double>&, double)> commit;
};

inline bool nm_close(double a, double b) { return std::fabs(a - b) <= 1e-13 * (std::fabs(a) + std::fabs(b)); }

inline NelderMeadResult nelder_mead(const BatchObjective& obj, std::vector<double> x0, const std::vector<double>& lb,
                                    const std::vector<double>& ub, const std::vector<double>& step, double ftol_abs,
                                    bool speculate = false, int max_eval = 0, int spec_batch = 4) {
    const int n_full = (int)x0.size();
    std::vector<int> dims;   // searched dimensions
    for (int i = 0; i < n_full; ++i) if (lb[i] != ub[i]) dims.push_back(i);
    const int n = (int)dims.size();
    NelderMeadResult R;
    auto full = [&](const std::vector<double>& y) {
        std::vector<double> x = x0;
        for (int k = 0; k < n; ++k) x[dims[k]] = y[k];
        return x;
    };
    auto eval = [&](const std::vector<std::vector<double>>& Y, bool record) {
        std::vector<std::vector<double>> X;
        for (const auto& y : Y) X.push_back(full(y));
        ++R.launches;
        if (record) R.evaluations += (int)Y.size();
        return obj.evaluate(X, record);
    };
    std::vector<double> l(n), u(n), s(n), y0(n);
    // a start value outside its bounds is moved onto the nearest bound before the simplex is built from it
    for (int k = 0; k < n; ++k) {
        l[k] = lb[dims[k]]; u[k] = ub[dims[k]]; s[k] = step[dims[k]];
        y0[k] = x0[dims[k]] = std::min(std::max(x0[dims[k]], l[k]), u[k]);
    }

    // initial simplex
    std::vector<std::vector<double>> P(n + 1, y0);
    for (int i = 0; i < n; ++i) {
        double& v = P[i + 1][i];
        v += s[i];
        if (v > u[i]) v = (u[i] - y0[i] > std::fabs(s[i]) * 0.1) ? u[i] : y0[i] - std::fabs(s[i]);
        if (v < l[i]) {
            if (y0[i] - l[i] > std::fabs(s[i]) * 0.1) v = l[i];
            else {
                v = y0[i] + std::fabs(s[i]);
                if (v > u[i]) v = 0.5 * ((u[i] - y0[i] > y0[i] - l[i] ? u[i] : l[i]) + y0[i]);
            }
        }
        if (nm_close(v, y0[i])) { R.x = x0; R.reason = "failure: a search dimension cannot be varied"; return R; }
    }
    std::vector<double> F = eval(P, true);
    double best_f = HUGE_VAL;
    std::vector<double> best_y = y0;
    auto track = [&](const std::vector<double>& y, double f) { if (f <= best_f) { best_f = f; best_y = y; } };
    for (int i = 0; i <= n; ++i) track(P[i], F[i]);

    // new = c + scale (c - old), clipped; false if it coincides with c or old
    auto reflect = [&](std::vector<double>& out, const std::vector<double>& c, double scale, const std::vector<double>& old) {
        bool eq_c = true, eq_old = true;
        std::vector<double> r(n);
        for (int i = 0; i < n; ++i) {
            double v = c[i] + scale * (c[i] - old[i]);
            if (v < l[i]) v = l[i];
            if (v > u[i]) v = u[i];
            eq_c = eq_c && nm_close(v, c[i]);
            eq_old = eq_old && nm_close(v, old[i]);
            r[i] = v;
        }
        out = r;
        return !(eq_c || eq_old);
    };
    const double alpha = 1, beta = 0.5, gamm = 2, delta = 0.5;
    auto finish = [&](const char* why) { R.x = full(best_y); R.f = best_f; R.reason = why; return R; };

    // ---- speculation: candidate points of this and the following iterations, breadth first ----
    // A hypothetical simplex: vertices with known values, and vertices accepted "in the future" whose value is unknown (NaN);
    // maybe_worst says whether such a vertex can be the worst one (a contraction point can, a reflection / expansion point is
    // below the second-highest value by the rule that accepted it).
    struct Hyp {
        std::vector<std::vector<double>> P;
        std::vector<double> F;
        std::vector<char> maybe_worst;
        int depth;
    };
    std::map<std::vector<double>, double> cache;
    auto collect = [&](const std::vector<std::vector<double>>& P0, const std::vector<double>& F0, const std::vector<double>& must,
                       std::vector<std::vector<double>>& out) {
        std::set<std::vector<double>> seen;
        auto add = [&](const std::vector<double>& y) {
            if ((int)out.size() >= spec_batch && !out.empty()) return;
            if (!seen.insert(y).second || cache.count(y)) return;
            out.push_back(y);
        };
        out.clear();
        seen.insert(must);
        out.push_back(must);
        std::deque<Hyp> q;
        q.push_back(Hyp{P0, F0, std::vector<char>(P0.size(), 0), 0});
        while (!q.empty() && (int)out.size() < spec_batch) {
            const Hyp h = q.front();
            q.pop_front();
            std::vector<int> worst;   // vertices that can be the highest one
            int km = -1;
            for (int i = 0; i <= n; ++i) {
                if (h.F[i] == h.F[i]) { if (km < 0 || h.F[i] >= h.F[km]) km = i; }
                else if (h.maybe_worst[i]) worst.push_back(i);
            }
            if (km >= 0) worst.insert(worst.begin(), km);
            for (int hi : worst) {
                std::vector<double> c(n, 0.0), xr, xe, xco, xci;
                for (int i = 0; i <= n; ++i) if (i != hi) for (int j = 0; j < n; ++j) c[j] += h.P[i][j];
                for (int j = 0; j < n; ++j) c[j] *= 1.0 / n;
                if (!reflect(xr, c, alpha, h.P[hi])) continue;
                const bool ok_e = reflect(xe, c, gamm, h.P[hi]), ok_co = reflect(xco, c, beta, h.P[hi]), ok_ci = reflect(xci, c, -beta, h.P[hi]);
                add(xr);
                if (ok_e) add(xe);
                if (ok_co) add(xco);
                if (ok_ci) add(xci);
                if (h.depth + 1 >= 4) continue;
                auto child = [&](const std::vector<double>& x, bool may_be_worst) {
                    Hyp k = h;
                    k.P[hi] = x;
                    k.F[hi] = std::nan("");
                    k.maybe_worst[hi] = may_be_worst;
                    k.depth = h.depth + 1;
                    q.push_back(k);
                };
                child(xr, false);
                if (ok_e) child(xe, false);
                if (ok_co) child(xco, true);
                if (ok_ci) child(xci, true);
            }
        }
    };
    // the value the sequential algorithm asks for next: from the cache, after one batched launch if it is not there yet
    auto need = [&](const std::vector<double>& y, const std::vector<std::vector<double>>& Pn, const std::vector<double>& Fn) {
        auto it = cache.find(y);
        if (it == cache.end()) {
            std::vector<std::vector<double>> pts;
            collect(Pn, Fn, y, pts);
            const std::vector<double> v = eval(pts, false);
            for (size_t k = 0; k < pts.size(); ++k) cache[pts[k]] = v[k];
            it = cache.find(y);
        }
        // a speculative batch reports nothing: a point that came back NaN is evaluated again on its own as a recorded evaluation,
        // so that the objective fails (or logs) exactly where the sequential algorithm would
        if (it->second != it->second) return eval({y}, true)[0];
        obj.commit(full(y), it->second);
        ++R.evaluations;
        return it->second;
    };
    for (;;) {
        // order: (f, index)
        int lo = 0, hi = 0;
        for (int i = 1; i <= n; ++i) {
            if (F[i] < F[lo]) lo = i;
            if (F[i] >= F[hi]) hi = i;
        }
        int second = lo;   // highest but one
        for (int i = 0; i <= n; ++i) if (i != hi && (F[i] > F[second] || (F[i] == F[second] && i > second))) second = i;
        if (std::fabs(F[lo] - F[hi]) < ftol_abs) return finish("ftol reached");
        if (max_eval > 0 && R.evaluations >= max_eval) return finish("maxeval reached");
        std::vector<double> c(n, 0.0);
        for (int i = 0; i <= n; ++i) if (i != hi) for (int j = 0; j < n; ++j) c[j] += P[i][j];
        for (int j = 0; j < n; ++j) c[j] *= 1.0 / n;

        std::vector<double> xr, xe;
        if (!reflect(xr, c, alpha, P[hi])) return finish("xtol reached");
        double fr, fe = 0;
        fr = speculate ? need(xr, P, F) : eval({xr}, true)[0];
        track(xr, fr);
        if (fr < F[lo]) {   // new best: try to expand
            if (!reflect(xe, c, gamm, P[hi])) return finish("xtol reached");
            fe = speculate ? need(xe, P, F) : eval({xe}, true)[0];
            track(xe, fe);
            if (fe >= fr) { P[hi] = xr; F[hi] = fr; } else { P[hi] = xe; F[hi] = fe; }
        } else if (fr < F[second]) {
            P[hi] = xr; F[hi] = fr;
        } else {            // new worst: contract
            const bool inside = F[hi] <= fr;
            std::vector<double> xc;
            double fc;
            if (!reflect(xc, c, inside ? -beta : beta, P[hi])) return finish("xtol reached");
            fc = speculate ? need(xc, P, F) : eval({xc}, true)[0];
            track(xc, fc);
            if (fc < fr && fc < F[hi]) { P[hi] = xc; F[hi] = fc; }
            else {          // shrink towards the best vertex
                std::vector<std::vector<double>> Q;
                std::vector<int> which;
                for (int i = 0; i <= n; ++i) {
                    if (i == lo) continue;
                    std::vector<double> q;
                    if (!reflect(q, P[lo], -delta, P[i])) return finish("xtol reached");
                    Q.push_back(q);
                    which.push_back(i);
                }
                const std::vector<double> v = eval(Q, true);
                for (size_t k = 0; k < Q.size(); ++k) { P[which[k]] = Q[k]; F[which[k]] = v[k]; track(Q[k], v[k]); }
            }
        }
    }
}

}  // namespace ggp

"""Synthetic lineage forests simulated from the model's own generative process (the reference ships no
generator; README.md:306 mentions one that is not in the repository).

Per tree: a root with x0 ~ N(log 2, 0.1^2), g0 ~ N(g_mean, (0.1 g_mean)^2) and (lambda, q) from their stationary
OU laws.  Between observations, lambda and q take exact OU transitions on `substeps` sub-intervals of dt, with
dx = lambda dt and dg = (e^x q - beta g) dt integrated on that sub-grid.  Every dt: x_obs = x + N(0, var_x),
g_obs = g + N(0, var_g) ("const") or N(0, var_g * g) ("scaled").  A cell divides after n_pts ~ U{lo..hi}
points; the gap mother-last -> daughter-first is one dt; x -> x - log 2 + N(0, var_dx); g -> g/2 + N(0, var_dg)
("gauss") or var_dg * Binomial(g / var_dg, 1/2) ("binomial"); both daughters are kept.  Times are k * dt.
Cells are stored tree-major (tree 0's cells in breadth-first order, then tree 1's, ...), so parents precede
daughters and a contiguous range of trees is a contiguous range of cells and time points.
"""
import numpy as np

from .forest import LineageData

# order: mean_lambda gamma_lambda var_lambda mean_q gamma_q var_q beta var_x var_g var_dx var_dg
# const noise / gauss division: the literals of reference tests.h:288-298 with var_lambda and var_q lowered
# (there the stationary sd of lambda is 2.2x its mean, which makes simulated lineages wander over e^+-10 in size)
PARAMS_CONST_GAUSS = np.array([0.01, 0.02, 2e-7, 10.0, 0.01, 0.02, 1e-3, 1e-3, 5000.0, 1e-3, 5000.0])
PARAMS_SCALED_BINOMIAL = np.array([2.31e-3, 4.36e-2, 9.75e-8, 9.03e1, 1.60e-2, 1.94e1, 4.95e-4, 3.35e-4, 8.26e-1,
                                   1.63e-4, 7.48e-1])   # reference example_data_set/parameter_file.txt:2-16


def simulate_forest(n_trees, n_generations, params=None, noise_model="const", division_model="gauss", dt=None,
                    pts_range=(17, 23), seed=20261018, substeps=8, n_segments=1, fp_auto=0.0):
    if params is None:
        params = PARAMS_CONST_GAUSS if noise_model == "const" else PARAMS_SCALED_BINOMIAL
    if dt is None:
        dt = 3.5 if noise_model == "const" else 15.0
    ml, gl, sl2, mq, gq, sq2, beta, var_x, var_g, var_dx, var_dg = [float(v) for v in params]
    rng = np.random.default_rng(seed)
    lo, hi = pts_range
    G = n_generations
    cells_per_tree = 2 ** G - 1
    n_cells = n_trees * cells_per_tree
    h = dt / substeps
    el, eq = np.exp(-gl * h), np.exp(-gq * h)
    sdl = np.sqrt(sl2 / (2 * gl) * (1 - el * el))
    sdq = np.sqrt(sq2 / (2 * gq) * (1 - eq * eq))
    g_mean = 2.0 * mq / (ml + beta) if noise_model == "const" else 2.0 * mq / (ml + beta) * 0.5

    npts_all = np.empty((n_trees, cells_per_tree), dtype=np.int64)
    t0_all = np.empty((n_trees, cells_per_tree), dtype=np.int64)   # time index of the first point
    gens = []
    # root states
    x = rng.normal(np.log(2.0), 0.1, size=(n_trees, 1))
    g = rng.normal(g_mean, 0.1 * g_mean, size=(n_trees, 1))
    lam = rng.normal(ml, np.sqrt(sl2 / (2 * gl)), size=(n_trees, 1))
    q = rng.normal(mq, np.sqrt(sq2 / (2 * gq)), size=(n_trees, 1))
    t0 = np.zeros((n_trees, 1), dtype=np.int64)
    for gen in range(G):
        w = 2 ** gen
        npts = rng.integers(lo, hi + 1, size=(n_trees, w))
        T = int(npts.max())
        xo = np.empty((n_trees, w, T))
        go = np.empty((n_trees, w, T))
        ex, eg, elam, eq_ = (np.empty((n_trees, w)) for _ in range(4))   # state one dt after the last point
        for k in range(T + 1):
            if k < T:
                xo[:, :, k] = x + rng.normal(0.0, np.sqrt(var_x), size=x.shape)
                sd_g = np.sqrt(var_g * np.maximum(g, 1e-300)) if noise_model == "scaled" else np.sqrt(var_g)
                go[:, :, k] = g + rng.normal(0.0, 1.0, size=g.shape) * sd_g
            if k > 0:
                done = npts == k
                ex[done], eg[done], elam[done], eq_[done] = x[done], g[done], lam[done], q[done]
            if k == T:
                break
            for _ in range(substeps):
                g = g + (np.exp(x) * q - beta * g) * h
                x = x + lam * h
                lam = ml + (lam - ml) * el + sdl * rng.normal(size=lam.shape)
                q = mq + (q - mq) * eq + sdq * rng.normal(size=q.shape)
        gens.append((npts, xo, go))
        b0 = w - 1
        npts_all[:, b0:b0 + w] = npts
        t0_all[:, b0:b0 + w] = t0
        if gen + 1 < G:
            # division: each mother (tree, j) -> daughters (tree, 2j), (tree, 2j+1)
            rep = lambda a: np.repeat(a, 2, axis=1)
            x = rep(ex) - np.log(2.0) + rng.normal(0.0, np.sqrt(var_dx), size=(n_trees, 2 * w))
            gm = rep(eg)
            if division_model == "binomial":
                n_mol = np.maximum(np.rint(gm / var_dg), 1).astype(np.int64)
                g = rng.binomial(n_mol, 0.5).astype(np.float64) * var_dg
            else:
                g = gm / 2.0 + rng.normal(0.0, np.sqrt(var_dg), size=gm.shape)
            lam, q = rep(elam), rep(eq_)
            t0 = rep(t0 + npts)
    # flatten: cell id = tree * cells_per_tree + breadth-first index
    npts_flat = npts_all.reshape(-1)
    cell_offset = np.concatenate([[0], np.cumsum(npts_flat)]).astype(np.int64)
    n_ctp = int(cell_offset[-1])
    time = np.empty(n_ctp)
    xs = np.empty(n_ctp)
    gs = np.empty(n_ctp)
    off2 = cell_offset[:-1].reshape(n_trees, cells_per_tree)
    for gen, (npts, xo, go) in enumerate(gens):
        w = 2 ** gen
        b0 = w - 1
        T = xo.shape[2]
        k = np.arange(T)[None, None, :]
        mask = k < npts[:, :, None]
        dest = (off2[:, b0:b0 + w, None] + k)[mask]
        xs[dest] = xo[mask]
        gs[dest] = go[mask]
        time[dest] = ((t0_all[:, b0:b0 + w, None] + k)[mask]) * dt
    b = np.arange(cells_per_tree)
    par_local = np.where(b > 0, (b - 1) // 2, -1)
    base = (np.arange(n_trees) * cells_per_tree)[:, None]
    parent = np.where(par_local[None, :] >= 0, base + par_local[None, :], -1).reshape(-1).astype(np.int32)
    segment = None
    if n_segments > 1:
        # piecewise parameters in time: split every tree's time span into n_segments equal parts
        t_end = (time.reshape(-1)[cell_offset[1:] - 1]).reshape(n_trees, cells_per_tree).max(axis=1)
        tree_of_ctp = np.repeat(np.repeat(np.arange(n_trees), cells_per_tree), npts_flat)
        segment = np.minimum((time / (t_end[tree_of_ctp] + dt) * n_segments).astype(np.int32), n_segments - 1)
    return LineageData(cell_offset=cell_offset, parent=parent, time=time, log_length=xs, fp=gs, segment=segment,
                       noise_model=noise_model, division_model=division_model, fp_auto=fp_auto)

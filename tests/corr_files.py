"""The reference's <prefix>_prediction.csv / <prefix>_joints.csv written from the CPU oracle's results (test
infrastructure): inputs of the correlation post-processing tests and of tools/make_correlation_golden.py."""
import os

import numpy as np

import gfp_gaussian_process_b200 as ggp
from oracle.oracle_py import Oracle

CASES = {
    "small": dict(trees=2, gens=3, pts=(3, 5), seed=51, n_data=12, noise="scaled", division="binomial"),
    "ragged": dict(trees=3, gens=2, pts=(1, 6), seed=52, n_data=8, noise="const", division="gauss"),
}


def g6(v):
    return "%g" % v


def write_case(case, outdir):
    """returns (joints file, prediction file, dt)"""
    P = ggp.PARAMS_CONST_GAUSS if case["noise"] == "const" else ggp.PARAMS_SCALED_BINOMIAL
    d = ggp.simulate_forest(case["trees"], case["gens"], params=P, noise_model=case["noise"], division_model=case["division"],
                            seed=case["seed"], pts_range=case["pts"])
    dt = float(np.min(np.diff(d.time[d.cell_offset[0]:d.cell_offset[1]]))) if d.cell_offset[1] > 1 else 1.0
    o = Oracle(d)
    pr = o.predictions([P])["prediction"]
    n, row, col, rec = o.joints(1e-10, 2000000)
    order = np.lexsort((col, row))
    row, col, rec = row[order], col[order], rec[order]
    ids = ["pos0.%d" % (c + 1) for c in range(d.n_cells)]   # not numeric: pandas (normalize_time branch of the script) must keep them strings
    pids = ["pos0.%d" % (p + 1) if p >= 0 else "pos0.0" for p in d.parent]
    cell_of = np.repeat(np.arange(d.n_cells), np.diff(d.cell_offset))
    head = "no,name,type,init,step,lower_bound,upper_bound,final\n" + "".join(
        "%d,%s,fixed,%s, , , ,\n" % (i, nme, g6(v)) for i, (nme, v) in enumerate(zip(ggp.PARAM_NAMES, P)))
    iu = [(a, b) for a in range(4) for b in range(a, 4)]
    pf = os.path.join(outdir, "forest_f_b_prediction.csv")
    with open(pf, "w") as f:
        f.write(head)
        f.write("\ncell_id,parent_id,time,log_length,fp,mean_x,mean_g,mean_l,mean_q,cov_xx,cov_xg,cov_xl,cov_xq,cov_gg,cov_gl,cov_gq,cov_ll,cov_lq,cov_qq\n")
        for k in range(d.n_ctp):
            c = cell_of[k]
            f.write(",".join([ids[c], pids[c], g6(d.time[k]), g6(d.log_length[k]), g6(d.fp[k])] + [g6(v) for v in pr[0][k]] +
                             [g6(pr[1][k][a][b]) for a, b in iu]) + "\n")
    jf = os.path.join(outdir, "forest_f_b_joints.csv")
    M = d.n_ctp
    with open(jf, "w") as f:
        f.write(head)
        f.write("\ncell_id,parent_id,time,")
        f.write("".join("%s_%s%s" % (ids[cell_of[k]], g6(d.time[k]), "," * (43 if k == M - 1 else 44)) for k in range(M)) + "\n")
        at = 0
        for r in range(M):
            c = cell_of[r]
            f.write("%s,%s,%s" % (ids[c], pids[c], g6(d.time[r])))
            nxt = 0
            while at < n and row[at] == r:
                f.write("," * (44 * (int(col[at]) - nxt)))
                f.write("".join("," + g6(v) for v in rec[at]))
                nxt = int(col[at]) + 1
                at += 1
            f.write("," * (44 * (M - nxt)) + "\n")
    return jf, pf, dt

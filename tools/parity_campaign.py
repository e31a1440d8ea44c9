"""Randomised parity campaign on the GPU box: CUDA path (through the C ABI) vs the oracle over random sizes, chunkings,
light curves and prior draws.  Prints a JSON summary (error percentiles, worst cases, error against condition number)
and writes it to gpurun_out/parity_campaign.json.  usage: python tools/parity_campaign.py [ncases] [seed]"""
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import gp_oracle as o  # noqa: E402
from gprotation_b200 import GPEngine  # noqa: E402
from gprotation_b200.synth import synthetic_lightcurve  # noqa: E402

import ctypes  # noqa: E402

_c = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "_build", "libgp_oracle.so"))
for _f in (_c.gp_oracle_lnlike_ld, _c.gp_oracle_lnlike_f64):
    _f.restype = ctypes.c_double
    _f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]


def lnlike_long_double(x, t, y, yerr, chunks, fn=None):
    """80-bit evaluation (oracle/gp_oracle_ld.c), summed over the chunks: the yardstick for both fp64 answers.
    fn = _c.gp_oracle_lnlike_f64: the unblocked fp64 Cholesky in george's operation order (a second fp64 opinion)."""
    fn = fn or _c.gp_oracle_lnlike_ld
    x = np.ascontiguousarray(x, dtype=np.float64)
    tot = 0.0
    for c in ([np.arange(len(t))] if chunks is None else chunks):
        tt, yy, ee = (np.ascontiguousarray(a[c], dtype=np.float64) for a in (t, y, yerr))
        tot += fn(x.ctypes.data, len(tt), tt.ctypes.data, yy.ctypes.data, ee.ctypes.data)
    return tot


ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 120
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
eng = GPEngine(0)
rows = []
t0 = time.time()
for case in range(ncases):
    n = int(rng.choice([rng.integers(1, 130), rng.integers(130, 700), rng.integers(700, 1600)]))
    kind = "gp" if (n <= 600 and rng.random() < 0.5) else "harmonic"
    t, y, yerr, truth = synthetic_lightcurve(int(rng.integers(1 << 30)), n, baseline=float(rng.uniform(0.3, 3.0)) * max(n, 40), kind=kind)
    if rng.random() < 0.3:
        yerr = yerr * rng.uniform(0.5, 20.0, size=n)          # heteroscedastic errors
    cs = None if (n < 4 or rng.random() < 0.5) else int(rng.integers(max(2, n // 8), n + 1))
    chunks = o.split_chunks(n, cs)
    eng.unregister_all()
    lc = eng.register_lc(t, y, yerr, chunks=chunks)
    th = o.sample_from_prior(5, seed=int(rng.integers(1 << 30)))
    th[0] = truth
    cl = int(rng.choice([0, 1, 2, 4, 64]))
    eng.set_cluster(cl)
    for exact in (False, True):
        got = eng.lnlike_batch(th, lc_id=lc, exact=exact)
        for k, x in enumerate(th):
            ref = o.lnlike(x, t, y, yerr, chunks)
            cond, terms = 0.0, 0.0
            for c in ([np.arange(n)] if chunks is None else chunks):          # worst chunk
                ev = np.linalg.eigvalsh(o.kernel_matrix(x, t[c], yerr[c]))
                cond = max(cond, float(ev[-1] / ev[0]) if ev[0] > 0 else float("inf"))
                if ev[0] > 0:                                                 # size of the terms lnlike is the (cancelling) sum of
                    terms += 0.5 * (abs(float(np.sum(np.log(ev)))) + len(c) * math.log(2 * math.pi))
            row = dict(n=n, chunksize=cs, kind=kind, cluster=cl, exact=exact, ref=float(ref), got=float(got[k]), cond=cond,
                       scale=max(abs(float(ref)), terms) if math.isfinite(ref) else 0.0)
            if math.isfinite(ref) and abs(got[k] - ref) > 1e-10 * abs(ref):
                ld = lnlike_long_double(x, t, y, yerr, chunks)
                f64 = lnlike_long_double(x, t, y, yerr, chunks, fn=_c.gp_oracle_lnlike_f64)
                row["gpu_vs_long_double"] = abs(got[k] - ld) / abs(ld)
                row["oracle_vs_long_double"] = abs(ref - ld) / abs(ld)
                row["unblocked_f64_vs_long_double"] = abs(f64 - ld) / abs(ld)
            rows.append(row)
eng.set_cluster(0)
fin = [r for r in rows if math.isfinite(r["ref"])]
rel = np.array([abs(r["got"] - r["ref"]) / abs(r["ref"]) for r in fin])
conds = np.array([r["cond"] for r in fin])
rel_sc = np.array([abs(r["got"] - r["ref"]) / r["scale"] for r in fin])
chk = [r for r in fin if "gpu_vs_long_double" in r]
tail = [r for r in chk if r["cond"] > 1e8]
tail_ratio = np.array([r["gpu_vs_long_double"] / max(r["oracle_vs_long_double"], r["unblocked_f64_vs_long_double"], 1e-300) for r in tail])
mism = [r for r in rows if math.isfinite(r["ref"]) != math.isfinite(r["got"])]
worst = sorted(fin, key=lambda r: -abs(r["got"] - r["ref"]) / abs(r["ref"]))[:8]
summary = {
    "cases": ncases, "evaluations": len(rows), "finite": len(fin), "minus_inf_in_both": len(rows) - len(fin) - len(mism),
    "finiteness_mismatches": len(mism), "seconds": time.time() - t0,
    "rel_err_percentiles": {str(p): float(np.percentile(rel, p)) for p in (50, 90, 99, 100)},
    "fraction_within_1e-9": float(np.mean(rel <= 1e-9)), "fraction_within_1e-10": float(np.mean(rel <= 1e-10)),
    "fraction_within_test_tolerance_1e-9_x_max(1,cond/1e7)": float(np.mean(rel <= 1e-9 * np.maximum(1.0, conds / 1e7))),
    "rel_err_over_eps_cond_percentiles": {str(p): float(np.percentile(rel / (1.1e-16 * conds), p)) for p in (50, 90, 99, 100)},
    "rule_of_tests/test_gpu_parity.py": {
        "flat_1e-9_vs_oracle": int(np.sum(rel <= 1e-9)),
        "else_within_3x_of_worse_cpu_fp64_from_80bit": int(sum(1 for r, e in zip(fin, rel) if e > 1e-9 and r["gpu_vs_long_double"] <= max(
            1e-9, 3 * max(r["oracle_vs_long_double"], r["unblocked_f64_vs_long_double"])))),
        "failing": [dict(n=r["n"], chunksize=r["chunksize"], cluster=r["cluster"], exact=r["exact"], cond=r["cond"], rel_err=float(e),
                         gpu_vs_long_double=r["gpu_vs_long_double"], oracle_vs_long_double=r["oracle_vs_long_double"],
                         unblocked_f64_vs_long_double=r["unblocked_f64_vs_long_double"])
                    for r, e in zip(fin, rel) if e > 1e-9 and not r["gpu_vs_long_double"] <= max(
                        1e-9, 3 * max(r["oracle_vs_long_double"], r["unblocked_f64_vs_long_double"]))]},
    "by_condition_number": {
        "cond<=1e8": {"evaluations": int(np.sum(conds <= 1e8)),
                      "max_rel_err": float(np.max(rel[conds <= 1e8])), "fraction_within_1e-9": float(np.mean(rel[conds <= 1e8] <= 1e-9)),
                      "max_err_over_max(|lnlike|,terms)": float(np.max(rel_sc[conds <= 1e8])),
                      "fraction_within_1e-9_of_max(|lnlike|,terms)": float(np.mean(rel_sc[conds <= 1e8] <= 1e-9))},
        "cond>1e8 (differing by more than 1e-10)": {
            "evaluations": len(tail),
            "gpu_distance_over_worse_of_two_cpu_fp64_distances_percentiles":
                {str(p): float(np.percentile(tail_ratio, p)) for p in (50, 90, 100)} if len(tail) else None,
            "fraction_within_3x": float(np.mean(tail_ratio <= 3.0)) if len(tail) else None}},
    "well_conditioned_subset(cond<=1e7)": {"evaluations": int(np.sum(conds <= 1e7)),
                                          "max_rel_err": float(np.max(rel[conds <= 1e7])) if np.any(conds <= 1e7) else None,
                                          "fraction_within_1e-9": float(np.mean(rel[conds <= 1e7] <= 1e-9)) if np.any(conds <= 1e7) else None},
    "against_80_bit_truth_where_the_two_differ_by_more_than_1e-10": {
        "evaluations": len(chk), "gpu_closer_than_oracle": int(sum(r["gpu_vs_long_double"] <= r["oracle_vs_long_double"] for r in chk)),
        "median_gpu": float(np.median([r["gpu_vs_long_double"] for r in chk])) if chk else None,
        "median_oracle": float(np.median([r["oracle_vs_long_double"] for r in chk])) if chk else None},
    "largest_in_units_of_eps_cond": [dict(n=r["n"], chunksize=r["chunksize"], cluster=r["cluster"], exact=r["exact"], cond=r["cond"],
                                          ref=r["ref"], rel_err=abs(r["got"] - r["ref"]) / abs(r["ref"]),
                                          units=abs(r["got"] - r["ref"]) / abs(r["ref"]) / (1.1e-16 * r["cond"]),
                                          gpu_vs_long_double=r.get("gpu_vs_long_double"), oracle_vs_long_double=r.get("oracle_vs_long_double"))
                                     for r in sorted(fin, key=lambda r: -abs(r["got"] - r["ref"]) / abs(r["ref"]) / r["cond"])[:6]],
    "worst": [dict(n=r["n"], chunksize=r["chunksize"], cluster=r["cluster"], exact=r["exact"], cond=r["cond"], ref=r["ref"],
                   rel_err=abs(r["got"] - r["ref"]) / abs(r["ref"]), gpu_vs_long_double=r.get("gpu_vs_long_double"),
                   oracle_vs_long_double=r.get("oracle_vs_long_double")) for r in worst],
}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(summary, open("gpurun_out/parity_campaign_seed%d.json" % seed, "w"), indent=1)
print(json.dumps(summary, indent=1))

"""Generates the committed golden fixtures from the reference itself.  Run in the BUILD container only
(needs /root/reference and oracle/_ref/libref_oracle.so):

    python tests/golden/make_golden.py

Outputs (small, committed):
  matern_gpytorch.json  the literal gpytorch tables of the reference's own tests, transcribed from
                        /root/reference/tests/test_radial.cc:212-489 (x grid, length scale, sigma,
                        15x15 Matern-5/2 and Matern-3/2 values), plus the scipy NLL known answer of
                        /root/reference/tests/test_evaluate.cc:34-63.
  ref_outputs.npz       outputs of the compiled reference (oracle/ref_shim) on seeded inputs drawn with
                        the reference's own generators (benchmarks/bench_utils.h:25-85): Gram matrices
                        for every menu covariance, fit / predict / nll / LOO / grouped CV, group
                        indexers, partition_triangular, sparse GP.
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.oracle import Ref  # noqa: E402

REF = "/root/reference"

PARAMS = {0: [2.0, 1.5], 1: [2.0, 1.5], 2: [3.0, 0.7], 3: [3.0, 0.7], 4: [1.3], 5: [0.2],
          6: [1.0, 1.0, 0.1], 7: [2.0, 1.5, 3.0, 0.7], 8: [2.0, 1.5, 3.0, 0.7, 0.1],
          9: [2.0, 1.5, 3.0, 0.7, 1.1, 0.9, 1.2, 0.3]}


def parse_tables():
    src = open(os.path.join(REF, "tests", "test_radial.cc")).read()

    def numbers(block):
        return [float(t) for t in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", block)]

    def after(name, count):
        i = src.index(name)
        j = src.index("{", i)
        depth, k = 0, j
        while True:
            depth += src[k] == "{"
            depth -= src[k] == "}"
            k += 1
            if depth == 0:
                break
        vals = numbers(src[j:k])
        assert len(vals) == count, (name, len(vals))
        return vals

    x = after("kOracleMaternX =", 15)
    m52 = after("kOracleMatern52Y{", 225)
    m32 = after("kOracleMatern32Y{", 225)
    ls = float(re.search(r"kMaternOracleLengthScale = ([\d.]+)", src).group(1))
    sg = float(re.search(r"kMaternOracleSigma = ([\d.]+)", src).group(1))
    return {"x": x, "length_scale": ls, "sigma": sg, "matern52": m52, "matern32": m32,
            "source": "tests/test_radial.cc:212-489 (python/gpytorch_covariance.py)"}


def main():
    golden = parse_tables()
    golden["nll_known_answer"] = {
        "x": [-1.0, 0.0, 1.0],
        "cov": [[1.0, 0.9, 0.8], [0.9, 1.0, 0.9], [0.8, 0.9, 1.0]],
        "value": 6.0946974293510134,
        "source": "tests/test_evaluate.cc:34-63 (scipy.stats.multivariate_normal)"}
    with open(os.path.join(HERE, "matern_gpytorch.json"), "w") as f:
        json.dump(golden, f, indent=1)

    out = {}
    x1 = Ref.random_features(96, 1, 0)
    x3 = Ref.random_features(80, 3, 0)
    out["x1"], out["x3"] = x1, x3
    for cid, p in PARAMS.items():
        out[f"gram1_{cid}"] = Ref.gram_sym(cid, p, x1)
        out[f"gram3_{cid}"] = Ref.gram_sym(cid, p, x3)
        out[f"cross3_{cid}"] = Ref.gram_cross(cid, p, x3[:30], x3[25:70])
    # exact GP on the reference's benchmark dataset shape
    xg = Ref.random_features(256, 1, 4)
    yg = Ref.random_targets(xg)
    xg3 = Ref.random_features(256, 3, 4)
    yg3 = Ref.random_targets(xg3)
    out["gp_x1"], out["gp_y1"], out["gp_x3"], out["gp_y3"] = xg, yg, xg3, yg3
    test1 = np.linspace(-0.5, 10.5, 23).reshape(-1, 1)
    test3 = Ref.random_features(17, 3, 11)
    out["gp_test1"], out["gp_test3"] = test1, test3
    for cid in (6, 8, 9):
        p = PARAMS[cid]
        for tag, x, y, t in (("1", xg, yg, test1), ("3", xg3, yg3, test3)):
            fit = Ref.gp_fit(cid, p, x, y)
            out[f"info{tag}_{cid}"] = fit["information"]
            out[f"nll{tag}_{cid}"] = np.array(Ref.gp_nll(cid, p, x, y)[0])
            out[f"mean{tag}_{cid}"] = Ref.gp_predict(cid, p, x, y, t, 0)[0]
            out[f"var{tag}_{cid}"] = Ref.gp_predict(cid, p, x, y, t, 1)[1]
            out[f"cov{tag}_{cid}"] = Ref.gp_predict(cid, p, x, y, t, 2)[2]
            m, v, _, s = Ref.gp_cv(cid, p, x, y, 0, 0.0, what=1, want_score=True)
            out[f"loo_mean{tag}_{cid}"], out[f"loo_var{tag}_{cid}"] = m, v
            out[f"loo_score{tag}_{cid}"] = np.array(s)
            keys, offsets, indices = Ref.group_indexers(x, 1, 8)
            m, v, _, s = Ref.gp_cv(cid, p, x, y, 1, 8.0, what=1, want_score=True)
            out[f"logo_mean{tag}_{cid}"], out[f"logo_var{tag}_{cid}"] = m, v
            out[f"logo_score{tag}_{cid}"] = np.array(s)
            _, _, j, _ = Ref.gp_cv(cid, p, x, y, 1, 8.0, what=2, group_sizes=np.diff(offsets))
            out[f"logo_joint{tag}_{cid}"] = j
    # with measurement variance on the targets (fit adds it, log_likelihood does not)
    yvar = 0.01 + 0.02 * (np.arange(256) % 5)
    out["gp_yvar"] = yvar
    out["info1_6_yvar"] = Ref.gp_fit(6, PARAMS[6], xg, yg, yvar=yvar)["information"]
    out["var1_6_yvar"] = Ref.gp_predict(6, PARAMS[6], xg, yg, test1, 1, yvar=yvar)[1]
    # integer contract
    keys, offsets, indices = Ref.group_indexers(xg, 1, 8)
    out["grp_keys"], out["grp_offsets"], out["grp_indices"] = keys, offsets, indices
    keys, offsets, indices = Ref.group_indexers(xg, 2, 3.7)
    out["grp2_keys"], out["grp2_offsets"], out["grp2_indices"] = keys, offsets, indices
    for n, k in ((100, 7), (1000, 8), (32768, 32), (5, 8)):
        out[f"ptri_{n}_{k}"] = Ref.partition_triangular(n, k)
    out["complement"] = Ref.indices_complement([3, 1, 7, 7, 12], 15)
    # LDLT wrapper on a bench-style PSD matrix (benchmarks/bench_utils.h:67-74)
    A = Ref.gram_sym(6, PARAMS[6], Ref.random_features(128, 1, 3))
    rhs = Ref.random_normal(128 * 3, 2).reshape(128, 3, order="F")
    l = Ref.ldlt(A, rhs=rhs, want_inverse_diagonal=True)
    out["ldlt_A"], out["ldlt_rhs"] = A, rhs
    out["ldlt_solve"], out["ldlt_sqrt_solve"] = l["solve"], l["sqrt_solve"]
    out["ldlt_logdet"] = np.array(l["logdet"])
    out["ldlt_inverse_diagonal"] = l["inverse_diagonal"]
    groups = [[0, 5, 9], [1], [100, 101, 102, 127], [64, 63]]
    out["ldlt_inverse_blocks"] = np.concatenate(
        [b.ravel(order="F") for b in Ref.inverse_blocks(A, groups)])
    # sparse GP (FITC and PITC groups)
    xs = Ref.random_features(400, 1, 5).ravel()
    ys = Ref.random_targets(xs)
    u = Ref.uniform_inducing_points(xs, 20)
    ts = np.linspace(0.2, 9.8, 13)
    out["sp_x"], out["sp_y"], out["sp_u"], out["sp_test"] = xs, ys, u, ts
    for tag, gk, ga in (("fitc", 0, 0.0), ("pitc", 2, 2.0)):
        r = Ref.sparse_gp(6, PARAMS[6], xs, ys, u, gk, ga, test=ts, what=2, want_ll=True)
        out[f"sp_{tag}_mean"], out[f"sp_{tag}_cov"] = r["mean"], r["cov"]
        out[f"sp_{tag}_ll"] = np.array(r["ll"])
        out[f"sp_{tag}_var"] = Ref.sparse_gp(6, PARAMS[6], xs, ys, u, gk, ga, test=ts,
                                             what=1)["var"]
    np.savez_compressed(os.path.join(HERE, "ref_outputs.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()

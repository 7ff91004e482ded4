"""GPU parity of the C++ trait layer: tests/cpp/trait_layer_check.cc uses the layer exactly as a
reference user writes code (gp_from_covariance, fit, predict().marginal(), log_likelihood,
cross_validate(), sparse_gp_from_covariance, DeviceLDLT) and dumps inputs + results; here they are
compared with the compiled reference (oracle/_ref) run on the very same inputs.
Tolerances: 1e-9 relative for means / information / likelihoods (north_star), 1e-9 relative to the
prior scale for variances (formed by cancellation in the reference as well), 4e-15 for Gram entries."""
import os
import subprocess

import numpy as np
import pytest

from oracle.oracle import Ref, Restate
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "trait_layer_check")
RTOL = 1e-9


def _parse(path):
    out = {}
    with open(path) as fh:
        while True:
            head = fh.readline()
            if not head:
                break
            key, count = head.split()
            out[key] = np.array([float(fh.readline()) for _ in range(int(count))])
    return out


@pytest.fixture(scope="module")
def dumped(tmp_path_factory):
    if not os.path.exists(EXE):
        pytest.fail("tests/cpp/trait_layer_check missing: run __graft_entry__.build()")
    path = str(tmp_path_factory.mktemp("cpp") / "out.txt")
    res = subprocess.run([EXE, "gpu", path], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout + res.stderr
    d = _parse(path)
    assert d["kernel_launches"][0] > 0
    return d


SINC = (6, [3.5, 5.7, 1.0])


def test_sinc_fit_predict_nll(dumped):
    d = dumped
    x, y = d["sinc.x"], d["sinc.y"]
    cid, p = SINC
    assert_close(d["sinc.information"], Ref.gp_fit(cid, p, x, y)["information"], RTOL, "information")
    mean, var, _ = Ref.gp_predict(cid, p, x, y, d["sinc.grid"], 1)
    assert_close(d["sinc.predict.mean"], mean, RTOL, "mean()")
    assert_close(d["sinc.predict.marginal.mean"], mean, RTOL, "marginal().mean")
    assert np.max(np.abs(d["sinc.predict.marginal.var"] - var)) <= 1e-9 * (5.7 ** 2 + 1.0)
    mean, _, cov = Ref.gp_predict(cid, p, x, y, d["sinc.few"], 2)
    assert_close(d["sinc.predict.joint.mean"], mean, RTOL, "joint().mean")
    assert np.max(np.abs(d["sinc.predict.joint.cov"].reshape(5, 5).T - cov)) <= 1e-9 * (5.7 ** 2 + 1.0)
    nll, _ = Ref.gp_nll(cid, p, x, y)
    assert abs(d["sinc.nll"][0] - nll) <= RTOL * abs(nll)
    nll, _ = Ref.gp_nll(cid, [2.0, 5.7, 0.5], x, y)
    assert abs(d["sinc.tuned.nll"][0] - nll) <= RTOL * abs(nll), "set_param_value must be live"


def test_sinc_leave_one_out(dumped):
    d = dumped
    x, y = d["sinc.x"], d["sinc.y"]
    cid, p = SINC
    mean, var, _, score = Ref.gp_cv(cid, p, x, y, 0, what=1, want_score=True)
    assert_close(d["sinc.loo.marginal.mean"], mean, RTOL, "LOO mean")
    assert_close(d["sinc.loo.mean"], mean, RTOL, "LOO mean()")
    assert_close(d["sinc.loo.marginal.var"], var, 1e-9, "LOO variance")
    # LeaveOneOutLikelihood<> = sum of per-point joint NLLs - prior ll (prior ll = 0 for these priors)
    assert abs(d["sinc.loo.likelihood"][0] - score) <= RTOL * abs(score)
    assert abs(d["sinc.loo.likelihood_marginal"][0] - score) <= RTOL * abs(score)
    rmse = np.sqrt((mean - y) ** 2)  # per-fold RMSE of one point = |error|; LeaveOneOutRMSE = their mean
    assert abs(d["sinc.loo.rmse"][0] - rmse.mean()) <= RTOL * rmse.mean()


def test_sinc_grouped_cross_validation(dumped):
    d = dumped
    x, y = d["sinc.x"], d["sinc.y"]
    cid, p = SINC
    sizes = d["sinc.cv.sizes"].astype(int)
    # integer contract: keys ascending, indices in encounter order, exactly the reference's indexer
    keys, offsets, indices = Ref.group_indexers(x.reshape(-1, 1), 1, 8.0)
    assert np.array_equal(d["sinc.cv.keys"].astype(np.int64), keys)
    assert np.array_equal(d["sinc.cv.indices"].astype(np.int64), indices)
    assert np.array_equal(sizes, np.diff(offsets))
    mean, var, _, _ = Ref.gp_cv(cid, p, x, y, 1, 8.0, what=1)
    assert_close(d["sinc.cv.marginal.mean"], mean, RTOL, "grouped mean")
    assert_close(d["sinc.cv.marginal.var"], var, 1e-9, "grouped variance")
    mean2, _, joint, score = Ref.gp_cv(cid, p, x, y, 1, 8.0, what=2, group_sizes=sizes, want_score=True)
    assert_close(d["sinc.cv.joint_blocks"], joint, 1e-9, "joint blocks")
    assert_close(d["sinc.cv.group_means"], mean2[indices], RTOL, "group means")
    # per-group scores = the reference's negative_log_likelihood(joint_g - truth_g)
    at, want = 0, []
    for g, k in enumerate(sizes):
        idx = indices[offsets[g]:offsets[g + 1]]
        cov = joint[at:at + k * k].reshape(k, k).T
        at += k * k
        want.append(Ref.nll_dense(mean2[idx] - y[idx], cov))
    assert_close(d["sinc.cv.scores"], np.array(want), 1e-9, "scores()")
    assert abs(d["sinc.cv.scores"].sum() - score) <= 1e-9 * abs(score)
    assert abs(d["sinc.cv.logo_likelihood"][0] - score) <= 1e-9 * abs(score)


def test_sinc_targets_with_measurement_variance(dumped):
    d = dumped
    x, y, yvar = d["sinc.x"], d["sinc.y"], d["sinc.yvar"]
    cid, p = SINC
    assert_close(d["sinc.noisy.information"], Ref.gp_fit(cid, p, x, y, yvar=yvar)["information"], RTOL)
    # held-out == brute-force refit (tests/test_cross_validation.cc:156-321), two of the groups
    keys, offsets, indices = Ref.group_indexers(x.reshape(-1, 1), 1, 8.0)
    for g in (0, len(keys) - 1):
        held = indices[offsets[g]:offsets[g + 1]]
        keep = np.setdiff1d(np.arange(len(x)), held)
        mean, _, cov = Ref.gp_predict(cid, p, x[keep], y[keep], x[held], 2, yvar=yvar[keep])
        # the held-out formula's covariance A_g^-1 is the refit's latent covariance PLUS the held
        # points' own measurement variance (they were inside K + diag(yvar) at the fit), and the
        # reference's metric then adds truth.covariance once more (prediction_metrics.hpp:112-118)
        want = Ref.nll_dense(mean - y[held], cov + 2.0 * np.diag(yvar[held]))
        assert abs(d["sinc.noisy.scores"][g] - want) <= 1e-9 * abs(want), (g, d["sinc.noisy.scores"][g], want)


def test_measurement_only(dumped):
    d = dumped
    x, y, t = d["meas.x"], d["meas.y"], d["meas.test"]
    p = [1.5, 2.0, 0.3]
    assert_close(d["meas.information"], Ref.gp_fit(10, p, x, y)["information"], RTOL, "information")
    mean, var, _ = Ref.gp_predict(10, p, x, y, t, 1)
    assert_close(d["meas.predict.marginal.mean"], mean, RTOL)
    assert np.max(np.abs(d["meas.predict.marginal.var"] - var)) <= 1e-9 * 4.0
    mean, _, cov = Ref.gp_predict(10, p, x, y, t, 2)
    assert np.max(np.abs(d["meas.predict.joint.cov"].reshape(5, 5).T - cov)) <= 1e-9 * 4.0
    mean, var, _ = Ref.gp_predict(10, p, x, y, t, 5)  # predict_with_measurement_noise
    assert_close(d["meas.predict_with_noise.marginal.mean"], mean, RTOL)
    assert np.max(np.abs(d["meas.predict_with_noise.marginal.var"] - var)) <= 1e-9 * 4.0
    # the noise term is present only between measurements
    assert np.all(d["meas.predict_with_noise.marginal.var"] - d["meas.predict.marginal.var"] > 0.08)
    nll, _ = Ref.gp_nll(10, p, x, y)
    assert abs(d["meas.nll"][0] - nll) <= RTOL * abs(nll)


def test_polynomial_sinc_example(dumped):
    """examples/sinc_example.cc:82-89: Polynomial<1> + SE + measurement_only(noise) through the trait layer."""
    d = dumped
    x, y, t = d["poly.x"], d["poly.y"], d["poly.test"]
    p = [3.0, 0.7, 3.5, 5.7, 0.4]
    assert_close(d["poly.gram"].reshape(12, 12).T, Ref.gram_sym(11, p, x[:12]), 1e-13, "gram")
    assert_close(d["poly.information"], Ref.gp_fit(11, p, x, y)["information"], RTOL, "information")
    mean, _, cov = Ref.gp_predict(11, p, x, y, t, 2)
    scale = np.max(np.abs(cov))
    assert_close(d["poly.predict.marginal.mean"], mean, RTOL)
    assert np.max(np.abs(d["poly.predict.joint.cov"].reshape(6, 6).T - cov)) <= 1e-9 * scale
    assert np.max(np.abs(d["poly.predict.marginal.var"] - np.diag(cov))) <= 1e-9 * scale
    mean, var, _ = Ref.gp_predict(11, p, x, y, t, 5)
    assert np.max(np.abs(d["poly.predict_with_noise.marginal.var"] - var)) <= 1e-9 * scale
    nll, _ = Ref.gp_nll(11, p, x, y)
    assert abs(d["poly.nll"][0] - nll) <= RTOL * abs(nll)
    m, _, _, _ = Ref.gp_cv(11, p, x, y, 0, 0.0, what=1, want_score=True)
    assert_close(d["poly.loo.mean"], m, RTOL, "LOO mean")
    # the linear prior extrapolates the trend: far outside the data the mean keeps rising
    assert d["poly.predict.marginal.mean"][5] > d["poly.predict.marginal.mean"][2] + 3.0


def test_3d_gram_and_gp(dumped):
    d = dumped
    x = d["v3.x"].reshape(-1, 3)
    n = len(x)
    p7, p8 = [2.0, 1.5, 3.0, 0.7], [2.0, 1.5, 3.0, 0.7, 0.1]
    K = Ref.gram_sym(7, p7, x)
    got = d["v3.gram"].reshape(n, n).T
    assert np.max(np.abs(got - K) / K) <= 4e-15
    assert np.array_equal(got, got.T)
    assert_close(d["v3.cross"].reshape(50, n).T, Ref.gram_cross(7, p7, x, x[:50]), 4e-15, "cross")
    assert_close(d["v3.diag"], Ref.gram_diag(7, p7, x), 4e-15, "diagonal")
    assert abs(d["v3.scalar"][0] - K[0, 1]) <= 4e-15 * K[0, 1]
    y = d["v3.y"]
    assert_close(d["v3.information"], Ref.gp_fit(8, p8, x, y)["information"], RTOL, "information")
    t = d["v3.test"].reshape(-1, 3)
    mean, _, cov = Ref.gp_predict(8, p8, x, y, t, 2)
    assert_close(d["v3.predict.joint.mean"], mean, RTOL)
    assert np.max(np.abs(d["v3.predict.joint.cov"].reshape(10, 10).T - cov)) <= 1e-9 * 2.75
    nll, _ = Ref.gp_nll(8, p8, x, y)
    assert abs(d["v3.nll"][0] - nll) <= RTOL * abs(nll)
    p9 = [2.0, 1.5, 3.0, 0.7, 1.5, 0.9, 1.1, 0.2]
    want = Ref.gram_sym(9, p9, x[:, 0])
    assert_close(d["sop.gram"].reshape(n, n).T, want, 4e-15, "sum of products")


def test_device_ldlt_surface(dumped):
    d = dumped
    n = 600
    K = d["ldlt.K"].reshape(n, n).T
    rhs = d["ldlt.rhs"].reshape(3, n).T
    want = Ref.ldlt(K, rhs=rhs, want_inverse_diagonal=True)
    assert_close(d["ldlt.solve"].reshape(3, n).T, want["solve"], RTOL, "solve")
    s = d["ldlt.sqrt_solve"].reshape(3, n).T
    assert_close(s.T @ s, rhs.T @ want["solve"], RTOL, "sqrt_solve identity")
    assert abs(d["ldlt.logdet"][0] - want["logdet"]) <= RTOL * abs(want["logdet"])
    assert_close(d["ldlt.inverse_diagonal"], want["inverse_diagonal"], RTOL, "inverse_diagonal")
    groups = [[0, 5, 9], [17], [400, 2, 3, 599]]
    blocks = Ref.inverse_blocks(K, groups)
    assert_close(d["ldlt.inverse_blocks"], np.concatenate([b.T.ravel() for b in blocks]), RTOL, "blocks")
    LD = d["ldlt.packed"].reshape(n, n).T
    L = np.tril(LD, -1) + np.eye(n)
    assert_close((L * np.diag(LD)) @ L.T, K, 1e-12, "L D L^T")


def test_sparse_gp(dumped):
    d = dumped
    x, y, t = d["sparse.x"], d["sparse.y"], d["sparse.test"]
    p = [1.0, 1.0, 0.1]
    u = Ref.uniform_inducing_points(x, 48)
    assert np.array_equal(d["sparse.inducing"], u)  # bit-exact: same accumulation as linspace
    want = Ref.sparse_gp(6, p, x, y, u, 0, test=t, what=1, want_ll=True)
    assert_close(d["sparse.fitc.marginal.mean"], want["mean"], RTOL, "FITC mean")
    assert np.max(np.abs(d["sparse.fitc.marginal.var"] - want["var"])) <= 1e-9
    assert abs(d["sparse.fitc.ll"][0] - want["ll"]) <= RTOL * abs(want["ll"])
    wj = Ref.sparse_gp(6, p, x, y, u, 0, test=t, what=2)
    assert np.max(np.abs(d["sparse.fitc.joint.cov"].reshape(19, 19).T - wj["cov"])) <= 1e-9
    want = Ref.sparse_gp(6, p, x, y, u, 2, 2.0, test=t, what=1, want_ll=True)
    assert_close(d["sparse.pitc.marginal.mean"], want["mean"], RTOL, "PITC mean")
    assert np.max(np.abs(d["sparse.pitc.marginal.var"] - want["var"])) <= 1e-9
    assert abs(d["sparse.pitc.ll"][0] - want["ll"]) <= RTOL * abs(want["ll"])


def test_sparse_gp_measurement_only(dumped):
    """SE + measurement_only(noise): the reference's standard sparse configuration
    (tests/lib/albatross/test/test_models.h:26-30), three covariance programs on the device."""
    d = dumped
    x, y, t = d["spmo.x"], d["spmo.y"], d["spmo.test"]
    p = [1.0, 1.0, 0.1]
    u = Ref.uniform_inducing_points(x, 40)
    assert np.array_equal(d["spmo.inducing"], u)
    want = Ref.sparse_gp(10, p, x, y, u, 0, test=t, what=1, want_ll=True)
    assert_close(d["spmo.fitc.marginal.mean"], want["mean"], RTOL, "FITC mean")
    assert np.max(np.abs(d["spmo.fitc.marginal.var"] - want["var"])) <= RTOL
    assert abs(d["spmo.fitc.ll"][0] - want["ll"]) <= RTOL * abs(want["ll"])
    wj = Ref.sparse_gp(10, p, x, y, u, 0, test=t, what=2)
    assert np.max(np.abs(d["spmo.fitc.joint.cov"].reshape(17, 17).T - wj["cov"])) <= RTOL
    want = Ref.sparse_gp(10, p, x, y, u, 2, 2.0, test=t, what=1, want_ll=True)
    assert_close(d["spmo.pitc.marginal.mean"], want["mean"], RTOL, "PITC mean")
    assert np.max(np.abs(d["spmo.pitc.marginal.var"] - want["var"])) <= RTOL
    assert abs(d["spmo.pitc.ll"][0] - want["ll"]) <= RTOL * abs(want["ll"])


def test_not_positive_definite_policy(dumped):
    """A singular K (duplicate points, no noise) yields NaN outputs, not an abort."""
    assert dumped["notpd.ll_is_nan"][0] == 1.0


def test_eigen_typed_build_matches_stand_in_types(dumped, tmp_path):
    """The same source compiled against the reference's vendored Eigen (Eigen::MatrixXd / VectorXd and
    Eigen-vector features instead of the stand-in types; tests/cpp/Makefile builds it where /root/reference
    exists and the binary travels to the GPU box) must produce the same numbers on the B200, bit for bit:
    the layer only moves column-major doubles through the C ABI."""
    exe = EXE + "_eigen"
    if not os.path.exists(exe):
        pytest.fail("tests/cpp/trait_layer_check_eigen missing: run __graft_entry__.build() where "
                    "/root/reference exists")
    path = str(tmp_path / "eigen.txt")
    res = subprocess.run([exe, "gpu", path], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout + res.stderr
    e = _parse(path)
    assert set(e) == set(dumped)
    for key, val in dumped.items():
        if key == "kernel_launches" or key in ("tune.tuned", "tune.before_after", "tune.gradient_at_1"):
            continue  # the simplex path and 1e-8 finite differences amplify the last bits of host-side sums
        # device results are bit-identical; scalars the layer reduces on the host (sums of per-group
        # scores) may differ in the last bits with Eigen's summation order
        assert val.shape == e[key].shape, key
        scale = float(np.max(np.abs(val[np.isfinite(val)]))) if np.any(np.isfinite(val)) else 1.0
        assert np.allclose(val, e[key], rtol=1e-13, atol=1e-13 * scale, equal_nan=True), key


def test_route_1b_real_reference_with_device_ldlt():
    """INTEGRATION.md route 1b, executed: tests/cpp/route1b_check.cc includes the REAL reference headers and
    instantiates the reference's own Fit<GPFit<CovarianceRepresentation, X>> (src/models/gp.hpp:42-78) and its
    generic _predict_impl (:305-366) with albatross_b200::DeviceLDLT; the binary compares every output with the
    stock SerializableLDLT model in-process at 1e-9."""
    exe = os.path.join(ROOT, "tests", "cpp", "route1b_check")
    if not os.path.exists(exe):
        pytest.fail("tests/cpp/route1b_check missing: run __graft_entry__.build() where /root/reference exists")
    res = subprocess.run([exe, "gpu"], capture_output=True, text=True, timeout=900)
    print(res.stdout)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "0 failure(s)" in res.stdout


def test_block_diagonal_ldlt_and_qr_concept(dumped):
    """BlockDiagonalLDLT (linalg/block_diagonal.hpp:96-218) and the QR concept (sparse_gp.hpp:72-89,
    linalg/qr_utils.hpp:18-53) of the layer: block solves against dense algebra on the same blocks, R^T R = B^T B
    and sqrt_solve = R^-T P^T rhs."""
    d = dumped
    K = d["bd.dense"].reshape(300, 300).T
    rhs = d["bd.rhs"].reshape(2, 300).T
    assert_close(d["bd.solve"].reshape(2, 300).T, np.linalg.solve(K, rhs), RTOL, "block solve")
    s = d["bd.sqrt_solve"].reshape(2, 300).T
    assert_close(s.T @ s, rhs.T @ np.linalg.solve(K, rhs), RTOL, "sqrt_solve Gram")
    sign, logdet = np.linalg.slogdet(K)
    assert sign > 0 and abs(d["bd.log_determinant"][0] - logdet) <= RTOL * abs(logdet)
    B = d["qr.B"].reshape(40, 340).T
    R = d["qr.R"].reshape(40, 40).T
    assert np.array_equal(R, np.triu(R)) and np.all(np.diag(R) > 0)
    assert_close(R.T @ R, B.T @ B, 1e-12, "R^T R = B^T B")
    _, Rnp = np.linalg.qr(B)
    Rnp = Rnp * np.sign(np.diag(Rnp))[:, None]       # Householder R up to row signs
    assert_close(R, Rnp, 1e-11, "R vs LAPACK")
    r2 = d["qr.rhs"].reshape(2, 40).T
    assert_close(d["qr.sqrt_solve"].reshape(2, 40).T, np.linalg.solve(R.T, r2), 1e-12, "R^-T rhs")


def test_tuner_objectives_match_the_reference(dumped):
    """the objective ModelTuner::tune() minimises (src/tune/tune.hpp:277-286) evaluated on the
    device for a batch of hyper-parameter candidates equals the reference's on the host at 1e-9 — both the
    default LeaveOneOutLikelihood (model_metrics.hpp:59-73) and the marginal likelihood; the loop itself
    improves the objective."""
    d = dumped
    x, y = d["tune.x"], d["tune.y"]
    cands = d["tune.candidates"].reshape(5, 3)
    for c in range(5):
        p = list(cands[c])
        _, _, _, score = Ref.gp_cv(6, p, x, y, 0, what=1, want_score=True)
        got = d["tune.candidate_objectives"][c]
        assert abs(got - score) <= RTOL * abs(score), (c, got, score)
        nll, _ = Ref.gp_nll(6, p, x, y)
        assert abs(d["tune.candidate_nll"][c] - nll) <= RTOL * abs(nll), (c, d["tune.candidate_nll"][c], nll)
    # forward differences of the reference objective (finite_difference.hpp:34-79) at candidate 1
    before, after, evals = d["tune.before_after"]
    assert after < before and evals <= 154
    tuned = list(d["tune.tuned"])
    _, _, _, score = Ref.gp_cv(6, tuned, x, y, 0, what=1, want_score=True)
    assert abs(after - score) <= RTOL * abs(score)


def test_update_through_the_layer(dumped):
    """fit_model.update(dataset) (core/fit_model.hpp:64-95 -> gp.hpp:386-414): partial fit + update == the
    reference's fit of everything (tests/test_gp.cc:182-219)."""
    d = dumped
    x, y, t = d["update.x"], d["update.y"], d["update.test"]
    p = [1.2, 1.1, 0.2]
    want = Ref.gp_fit(10, p, x, y)["information"]
    assert_close(d["update.information"], want, RTOL, "information after update")
    assert_close(d["update.full_information"], want, RTOL, "information of the full fit")
    mean, var, _ = Ref.gp_predict(10, p, x, y, t, 1)
    assert_close(d["update.marginal.mean"], mean, RTOL, "mean after update")
    assert np.max(np.abs(d["update.marginal.var"] - var)) <= RTOL * 1.3

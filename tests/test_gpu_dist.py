"""GPU parity of the distributed entry points on ONE GPU (world = 1 exercises the whole blocked
block-column algorithm without collectives; the sharded CV units are summed on the host).  The
multi-rank runs proper are tools/dist_check.py under torchrun (gpurun --gpus 2/4/8)."""
import numpy as np
import pytest

from albatross_b200 import capi
from albatross_b200.capi import MARGINAL, MEAN
from oracle.oracle import Restate, group_keys
from tests.helpers import assert_close, features, prog, targets

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.mark.parametrize("n,nb,dim,cid", [(100, 128, 1, 6), (1000, 128, 3, 8), (1500, 256, 3, 8),
                                          (2049, 512, 1, 6), (3000, 1024, 3, 9)])
def test_dist_fit_world1_matches_oracle(handle, n, nb, dim, cid):
    ops, pp = prog(cid)
    x = features(n, dim, n)
    y = targets(x)
    want = Restate.gp_fit(ops, pp, x, y)
    want_nll = Restate.gp_nll(ops, pp, x, y)
    f, info, nll = handle.dist_gp_fit(ops, pp, x, y, nb=nb)
    assert_close(info, want["information"], RTOL, "information")
    assert abs(nll - want_nll) <= RTOL * abs(want_nll)
    f.free()


def test_dist_fit_world1_with_measurement_variance(handle):
    ops, pp = prog(8)
    x = features(700, 3, 4)
    y = targets(x)
    yvar = 0.02 + 0.1 * np.random.default_rng(0).uniform(size=len(y))
    want = Restate.gp_fit(ops, pp, x, y, yvar=yvar)
    f, info, _ = handle.dist_gp_fit(ops, pp, x, y, yvar=yvar, nb=256)
    assert_close(info, want["information"], RTOL)
    f.free()


def test_dist_fit_matches_single_gpu_path(handle):
    """Same device, two algorithms (recursive in-place vs block-column): information and nll agree
    far inside the parity budget at a size the oracle would need minutes for."""
    ops, pp = prog(8)
    n = 6000
    x = features(n, 3, 11)
    y = targets(x)
    f1, info1 = handle.gp_fit(ops, pp, x, y)
    nll1 = handle.gp_nll(ops, pp, x, y)
    f2, info2, nll2 = handle.dist_gp_fit(ops, pp, x, y, nb=1024)
    assert_close(info2, info1, 1e-10)
    assert abs(nll1 - nll2) <= 1e-12 * abs(nll1)
    f1.free()
    f2.free()


def test_dist_not_positive_definite(handle):
    ops, pp = prog(7)  # SE + Matern52 without noise: duplicate points make K singular
    x = np.repeat(features(200, 3, 1), 2, axis=0)
    y = targets(x)
    with pytest.raises(capi.AbError) as e:
        handle.dist_gp_fit(ops, pp, x, y, nb=128)
    assert e.value.status == 4


def test_dist_gram_rows_world1(handle):
    ops, pp = prog(7)
    x = features(333, 3, 2)
    r0, K = handle.dist_gram_rows(ops, pp, x)
    assert r0 == 0 and K.shape == (333, 333)
    assert_close(K.download(), Restate.gram_sym(ops, pp, x), 1e-14)
    K.free()


@pytest.mark.parametrize("grouper", [(1, 8.0), (2, 1.0)])
def test_cv_shards_sum_to_full(handle, grouper):
    ops, pp = prog(6)
    n = 900
    x = features(n, 1, 27).ravel()
    y = targets(x)
    keys = group_keys(x, *grouper)
    _, offsets, indices = capi.group_indexers(keys)
    f, info = handle.gp_fit(ops, pp, x, y)
    mean, var, _, score = handle.gp_cv(f, y, info, offsets, indices, MARGINAL, want_score=True)
    want_mean, want_var, _, want_score = Restate.gp_cv(ops, pp, x, y, keys, what=1, want_score=True)
    assert_close(mean, want_mean, RTOL)
    for nshards in (2, 3):
        ms, vs, ss = np.zeros(n), np.zeros(n), 0.0
        for s in range(nshards):
            m_, v_, s_ = handle.gp_cv_shard(f, y, info, offsets, indices, MARGINAL, s, nshards,
                                            want_score=True)
            ms += m_
            vs += v_
            ss += s_
        assert_close(ms, want_mean, RTOL, "sharded means")
        assert_close(vs, want_var, 1e-9, "sharded variances")
        assert abs(ss - want_score) <= 1e-9 * abs(want_score)
    f.free()


def test_loo_shards_sum_to_full(handle):
    """Pure leave-one-out: the inverse diagonal is computed in 2048-column chunks per shard."""
    ops, pp = prog(6)
    n = 5000
    x = features(n, 1, 3).ravel()
    y = targets(x)
    keys = np.arange(n, dtype=np.int64)
    _, offsets, indices = capi.group_indexers(keys)
    f, info = handle.gp_fit(ops, pp, x, y)
    mean, var, _, score = handle.gp_cv(f, y, info, offsets, indices, MARGINAL, want_score=True)
    ms, vs, ss = np.zeros(n), np.zeros(n), 0.0
    for s in range(2):
        m_, v_, s_ = handle.gp_cv_shard(f, y, info, offsets, indices, MARGINAL, s, 2,
                                        want_score=True)
        ms += m_
        vs += v_
        ss += s_
    assert_close(ms, mean, 1e-10)
    assert_close(vs, var, 1e-10)
    assert abs(ss - score) <= 1e-10 * abs(score)
    f.free()


def test_world1_broadcast_and_breakdown(handle):
    """ab_dist_factor_broadcast is the identity on a one-rank group; ab_dist_fit_breakdown reports the per-step
    accounting of the most recent distributed fit (the decoupled panel pipeline also runs without NCCL)."""
    ops, pp = prog(8)
    x = features(2300, 3, 4)
    y = targets(x)
    f, info = handle.gp_fit(ops, pp, x, y)
    g = handle.dist_factor_broadcast(f, root=0)
    assert g is f
    df, dinfo, nll = handle.dist_gp_fit(ops, pp, x, y, nb=256)
    wait_ms, panel_ms, steps = handle.dist_fit_breakdown()
    assert steps == 9 and wait_ms >= 0.0 and panel_ms > 0.0
    assert_close(dinfo, info, 1e-10, "distributed (world 1, pipeline) vs single-GPU information")
    df.free()


@pytest.mark.parametrize("schedule", ["lookahead1", "pipeline", "pipeline_unpaired", "pipeline_nbuf8"])
def test_dist_schedules_agree(handle, schedule, monkeypatch):
    """Every schedule of the distributed factorisation (AB_DIST_SCHEDULE, AB_DIST_PAIR, AB_DIST_NBUF) gives the
    oracle's answer: round-1 style look-ahead, the panel pipeline with paired (default) and single-panel bulk
    updates, and a deeper ring of packed-panel buffers."""
    if schedule == "lookahead1":
        monkeypatch.setenv("AB_DIST_SCHEDULE", "lookahead1")
    elif schedule == "pipeline_unpaired":
        monkeypatch.setenv("AB_DIST_PAIR", "0")
    elif schedule == "pipeline_nbuf8":
        monkeypatch.setenv("AB_DIST_NBUF", "8")
    ops, pp = prog(6)
    x = features(1700, 1, 9).ravel()
    y = targets(x)
    df, info, nll = handle.dist_gp_fit(ops, pp, x, y, nb=128)
    assert_close(info, Restate.gp_fit(ops, pp, x, y)["information"], 1e-9, f"{schedule} information")
    want = Restate.gp_nll(ops, pp, x, y)
    assert abs(nll - want) <= 1e-9 * abs(want)
    df.free()

"""TEST INFRASTRUCTURE ONLY -- ctypes front-ends for the two CPU checkers.

* ``Ref``      -> ``oracle/_ref/libref_oracle.so``: the REAL reference (swift-nav/albatross headers
  compiled in place by ``oracle/Makefile``; see ``oracle/ref_shim/*.cc`` for the entry points and the
  reference file:line each one drives).
* ``Restate``  -> ``oracle/liboracle_restate.so``: the plain-C restatement of the same algorithms
  (``oracle/restate.c``), usable for arbitrary covariance programs.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this module.  The product (``albatross_b200/``) must never do so.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libref_oracle.so")
RESTATE_SO = os.path.join(HERE, "liboracle_restate.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)


def _d(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] or a.flags["F_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def _i(a):
    if a is None:
        return None
    assert a.dtype == np.int64
    return a.ctypes.data_as(_ip)


def _feats(x):
    """Features as a C-contiguous (n, dim) float64 array (AoS, as std::vector<X>)."""
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    if x.ndim == 1:
        x = x.reshape(-1, 1)
    return x


def build(target="all"):
    subprocess.check_call(["make", "-s", "-C", HERE, target])


# Covariance menu of the compiled reference (oracle/ref_shim/ref_common.h); value = the equivalent
# postfix program understood by the restatement and by the device (include/albatross_b200.h).
SE, EXP, M32, M52, CONST, NOISE, SUM, PROD, POLY = 1, 2, 3, 4, 5, 6, 7, 8, 9


def menu_program(cov_id, p):
    """(ops, params) postfix program equivalent to reference menu entry ``cov_id``."""
    p = list(map(float, p))
    if cov_id in (0, 1, 2, 3):
        return [(SE, EXP, M32, M52)[cov_id]], [p[0], p[1]]
    if cov_id == 4:
        return [CONST], [p[0], 0.0]
    if cov_id == 5:
        return [NOISE], [p[0], 0.0]
    if cov_id == 6:
        return [SE, NOISE, SUM], [p[0], p[1], p[2], 0.0, 0.0, 0.0]
    if cov_id == 7:
        return [SE, M52, SUM], [p[0], p[1], p[2], p[3], 0.0, 0.0]
    if cov_id == 8:
        return [SE, M52, SUM, NOISE, SUM], [p[0], p[1], p[2], p[3], 0, 0, p[4], 0, 0, 0]
    if cov_id == 9:
        # SE*M32 + EXP*CONST + NOISE
        return (
            [SE, M32, PROD, EXP, CONST, PROD, SUM, NOISE, SUM],
            [p[0], p[1], p[2], p[3], 0, 0, p[4], p[5], p[6], 0, 0, 0, 0, 0, p[7], 0, 0, 0],
        )
    if cov_id == 10:
        # SE + measurement_only(NOISE), as seen between two Measurement<> features (the fit)
        return [SE, NOISE, SUM], [p[0], p[1], p[2], 0.0, 0.0, 0.0]
    if cov_id == 11:
        # Polynomial<1> + SE + measurement_only(NOISE) between two Measurement<> features:
        # ((t0 + t1) + SE) + NOISE, each polynomial term (sigma, degree)
        return ([POLY, POLY, SUM, SE, SUM, NOISE, SUM],
                [p[0], 0.0, p[1], 1.0, 0, 0, p[2], p[3], 0, 0, p[4], 0.0, 0, 0])
    raise ValueError(cov_id)


def menu_program_plain(cov_id, p):
    """Program of menu entry ``cov_id`` when at least one side is NOT a Measurement<> (K_fu, K_uu and
    the predictions of the sparse GP; cross / prior of the exact GP): MeasurementOnly terms vanish
    (measurement.hpp:70-114).  Only entry 10 differs from menu_program."""
    if cov_id == 10:
        p = list(map(float, p))
        return [SE], [p[0], p[1]]
    if cov_id == 11:
        p = list(map(float, p))
        return [POLY, POLY, SUM, SE, SUM], [p[0], 0.0, p[1], 1.0, 0, 0, p[2], p[3], 0, 0]
    return menu_program(cov_id, p)


class Ref:
    """The compiled reference.  ``Ref.available()`` is False when the .so was not built/shipped."""

    _lib = None

    @classmethod
    def available(cls):
        return os.path.exists(REF_SO)

    @classmethod
    def lib(cls):
        if cls._lib is None:
            lib = C.CDLL(REF_SO)
            lib.ref_cov_scalar.restype = C.c_double
            lib.ref_cov_scalar.argtypes = [C.c_int, _dp, C.c_double, C.c_double]
            lib.ref_nll_dense.restype = C.c_double
            lib.ref_group_indexers.restype = C.c_int64
            lib.ref_partition_triangular.restype = C.c_int64
            lib.ref_indices_complement.restype = C.c_int64
            cls._lib = lib
        return cls._lib

    # -- generators -------------------------------------------------------------------------
    @classmethod
    def random_features(cls, n, dim=1, seed=0):
        out = np.empty((n, dim))
        cls.lib().ref_random_features(C.c_int64(n), C.c_int(dim), C.c_uint32(seed), _d(out))
        return out

    @classmethod
    def random_targets(cls, feats):
        x = _feats(feats)
        out = np.empty(x.shape[0])
        cls.lib().ref_random_targets(_d(x), C.c_int64(x.shape[0]), C.c_int(x.shape[1]), _d(out))
        return out

    @classmethod
    def random_normal(cls, n, seed=1):
        out = np.empty(n)
        cls.lib().ref_random_normal(C.c_int64(n), C.c_uint32(seed), _d(out))
        return out

    # -- gram ---------------------------------------------------------------------------------
    @classmethod
    def gram_sym(cls, cov_id, params, feats, as_meas=False, nthreads=1):
        x = _feats(feats)
        n, dim = x.shape
        p = np.asarray(params, dtype=np.float64)
        out = np.empty((n, n), order="F")
        rc = cls.lib().ref_gram_sym(C.c_int(cov_id), _d(p), _d(x), C.c_int64(n), C.c_int(dim),
                                    C.c_int(int(as_meas)), C.c_int(nthreads), _d(out))
        assert rc == 0
        return out

    @classmethod
    def gram_cross(cls, cov_id, params, fx, fy, nthreads=1):
        x, y = _feats(fx), _feats(fy)
        assert x.shape[1] == y.shape[1]
        p = np.asarray(params, dtype=np.float64)
        out = np.empty((x.shape[0], y.shape[0]), order="F")
        rc = cls.lib().ref_gram_cross(C.c_int(cov_id), _d(p), _d(x), C.c_int64(x.shape[0]), _d(y),
                                      C.c_int64(y.shape[0]), C.c_int(x.shape[1]),
                                      C.c_int(nthreads), _d(out))
        assert rc == 0
        return out

    @classmethod
    def gram_diag(cls, cov_id, params, feats):
        x = _feats(feats)
        p = np.asarray(params, dtype=np.float64)
        out = np.empty(x.shape[0])
        rc = cls.lib().ref_gram_diag(C.c_int(cov_id), _d(p), _d(x), C.c_int64(x.shape[0]),
                                     C.c_int(x.shape[1]), _d(out))
        assert rc == 0
        return out

    @classmethod
    def cov_scalar(cls, cov_id, params, x, y):
        p = np.asarray(params, dtype=np.float64)
        return cls.lib().ref_cov_scalar(C.c_int(cov_id), _d(p), C.c_double(x), C.c_double(y))

    # -- LDLT wrapper -------------------------------------------------------------------------
    @classmethod
    def ldlt(cls, A, rhs=None, want_inverse_diagonal=False):
        A = np.asfortranarray(A, dtype=np.float64)
        n = A.shape[0]
        out = {
            "ldlt": np.empty((n, n), order="F"),
            "transpositions": np.empty(n, dtype=np.int64),
            "D": np.empty(n),
        }
        logdet = C.c_double()
        is_pd = C.c_int()
        nrhs = 0
        rhs_f = None
        if rhs is not None:
            rhs_f = np.asfortranarray(np.asarray(rhs, dtype=np.float64).reshape(n, -1))
            nrhs = rhs_f.shape[1]
            out["solve"] = np.empty((n, nrhs), order="F")
            out["sqrt_solve"] = np.empty((n, nrhs), order="F")
        if want_inverse_diagonal:
            out["inverse_diagonal"] = np.empty(n)
        cls.lib().ref_ldlt(_d(A), C.c_int64(n), _d(out["ldlt"]), _i(out["transpositions"]),
                           _d(out["D"]), C.byref(logdet), C.byref(is_pd), _d(rhs_f),
                           C.c_int64(nrhs), _d(out.get("solve")), _d(out.get("sqrt_solve")),
                           _d(out.get("inverse_diagonal")))
        out["logdet"] = logdet.value
        out["is_pd"] = bool(is_pd.value)
        return out

    @classmethod
    def inverse_blocks(cls, A, groups, nthreads=1):
        A = np.asfortranarray(A, dtype=np.float64)
        n = A.shape[0]
        indices = np.concatenate([np.asarray(g, dtype=np.int64) for g in groups])
        offsets = np.zeros(len(groups) + 1, dtype=np.int64)
        offsets[1:] = np.cumsum([len(g) for g in groups])
        out = np.empty(int(sum(len(g) ** 2 for g in groups)))
        cls.lib().ref_ldlt_inverse_blocks(_d(A), C.c_int64(n), _i(indices), _i(offsets),
                                          C.c_int64(len(groups)), C.c_int(nthreads), _d(out))
        blocks, c = [], 0
        for g in groups:
            k = len(g)
            blocks.append(out[c:c + k * k].reshape(k, k, order="F"))
            c += k * k
        return blocks

    # -- integer contract ---------------------------------------------------------------------
    @classmethod
    def group_indexers(cls, feats, grouper_kind, grouper_arg=0.0):
        x = _feats(feats)
        n, dim = x.shape
        keys = np.empty(n, dtype=np.int64)
        offsets = np.empty(n + 1, dtype=np.int64)
        indices = np.empty(n, dtype=np.int64)
        g = cls.lib().ref_group_indexers(_d(x), C.c_int64(n), C.c_int(dim), C.c_int(grouper_kind),
                                         C.c_double(grouper_arg), _i(keys), _i(offsets),
                                         _i(indices))
        return keys[:g].copy(), offsets[:g + 1].copy(), indices

    @classmethod
    def partition_triangular(cls, n, count):
        out = np.empty(2 * count, dtype=np.int64)
        k = cls.lib().ref_partition_triangular(C.c_int64(n), C.c_int64(count), _i(out))
        return out[:2 * k].reshape(k, 2)

    @classmethod
    def indices_complement(cls, indices, n):
        idx = np.asarray(indices, dtype=np.int64)
        out = np.empty(n, dtype=np.int64)
        k = cls.lib().ref_indices_complement(_i(idx), C.c_int64(len(idx)), C.c_int64(n), _i(out))
        return out[:k].copy()

    @classmethod
    def nll_dense(cls, deviation, cov):
        d = np.ascontiguousarray(deviation, dtype=np.float64)
        c = np.asfortranarray(cov, dtype=np.float64)
        return cls.lib().ref_nll_dense(_d(d), _d(c), C.c_int64(len(d)))

    # -- exact GP -----------------------------------------------------------------------------
    @classmethod
    def gp_fit(cls, cov_id, params, feats, y, yvar=None, nthreads=1, want_factor=False):
        x = _feats(feats)
        n, dim = x.shape
        p = np.asarray(params, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        yv = None if yvar is None else np.ascontiguousarray(yvar, dtype=np.float64)
        out = {"information": np.empty(n)}
        if want_factor:
            out["ldlt"] = np.empty((n, n), order="F")
            out["transpositions"] = np.empty(n, dtype=np.int64)
            out["D"] = np.empty(n)
        rc = cls.lib().ref_gp_fit(C.c_int(cov_id), _d(p), _d(x), C.c_int64(n), C.c_int(dim), _d(y),
                                  _d(yv), C.c_int(nthreads), _d(out["information"]),
                                  _d(out.get("ldlt")), _i(out.get("transpositions")),
                                  _d(out.get("D")))
        assert rc == 0
        return out

    @classmethod
    def gp_predict(cls, cov_id, params, feats, y, test, what, yvar=None):
        x, t = _feats(feats), _feats(test)
        n, dim = x.shape
        pn = t.shape[0]
        p = np.asarray(params, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        yv = None if yvar is None else np.ascontiguousarray(yvar, dtype=np.float64)
        mean = np.empty(pn)
        var = np.empty(pn) if what in (1, 5) else None  # 5 = predict_with_measurement_noise marginal
        cov = np.empty((pn, pn), order="F") if what == 2 else None
        rc = cls.lib().ref_gp_predict(C.c_int(cov_id), _d(p), _d(x), C.c_int64(n), C.c_int(dim),
                                      _d(y), _d(yv), _d(t), C.c_int64(pn), C.c_int(what), _d(mean),
                                      _d(var), _d(cov))
        assert rc == 0
        return mean, var, cov

    @classmethod
    def gp_nll(cls, cov_id, params, feats, y):
        x = _feats(feats)
        n, dim = x.shape
        p = np.asarray(params, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        nll, prior = C.c_double(), C.c_double()
        rc = cls.lib().ref_gp_nll(C.c_int(cov_id), _d(p), _d(x), C.c_int64(n), C.c_int(dim), _d(y),
                                  C.byref(nll), C.byref(prior))
        assert rc == 0
        return nll.value, prior.value

    @classmethod
    def gp_cv(cls, cov_id, params, feats, y, grouper_kind, grouper_arg=0.0, what=1, nthreads=1,
              group_sizes=None, want_score=False):
        x = _feats(feats)
        n, dim = x.shape
        p = np.asarray(params, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        mean = np.empty(n)
        var = np.empty(n) if what == 1 else None
        joint = None
        if what == 2:
            assert group_sizes is not None
            joint = np.empty(int(sum(int(k) ** 2 for k in group_sizes)))
        score = C.c_double()
        rc = cls.lib().ref_gp_cv(C.c_int(cov_id), _d(p), _d(x), C.c_int64(n), C.c_int(dim), _d(y),
                                 C.c_int(grouper_kind), C.c_double(grouper_arg), C.c_int(nthreads),
                                 C.c_int(what), _d(mean), _d(var), _d(joint),
                                 C.byref(score) if want_score else None)
        assert rc == 0
        return mean, var, joint, (score.value if want_score else None)

    # -- sparse GP ----------------------------------------------------------------------------
    @classmethod
    def uniform_inducing_points(cls, feats, m):
        x = np.ascontiguousarray(feats, dtype=np.float64).ravel()
        out = np.empty(m)
        cls.lib().ref_uniform_inducing_points(_d(x), C.c_int64(len(x)), C.c_int64(m), _d(out))
        return out

    @classmethod
    def sparse_gp(cls, cov_id, params, feats, y, inducing, grouper_kind, grouper_arg=0.0,
                  test=None, what=-1, yvar=None, measurement_nugget=-1.0, inducing_nugget=-1.0,
                  want_ll=False):
        x = np.ascontiguousarray(feats, dtype=np.float64).ravel()
        n = len(x)
        u = np.ascontiguousarray(inducing, dtype=np.float64)
        m = len(u)
        p = np.asarray(params, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        yv = None if yvar is None else np.ascontiguousarray(yvar, dtype=np.float64)
        t = None if test is None else np.ascontiguousarray(test, dtype=np.float64).ravel()
        pn = 0 if t is None else len(t)
        out = {"information": np.empty(m), "R": np.empty((m, m), order="F"),
               "perm": np.empty(m, dtype=np.int64)}
        rank = C.c_int64()
        ll = C.c_double()
        if what >= 0:
            out["mean"] = np.empty(pn)
        if what == 1:
            out["var"] = np.empty(pn)
        if what == 2:
            out["cov"] = np.empty((pn, pn), order="F")
        rc = cls.lib().ref_sparse_gp(
            C.c_int(cov_id), _d(p), _d(x), C.c_int64(n), _d(y), _d(yv), _d(u), C.c_int64(m),
            C.c_int(grouper_kind), C.c_double(grouper_arg), C.c_double(measurement_nugget),
            C.c_double(inducing_nugget), _d(t), C.c_int64(pn), C.c_int(what),
            _d(out["information"]), _d(out["R"]), _i(out["perm"]), C.byref(rank),
            _d(out.get("mean")), _d(out.get("var")), _d(out.get("cov")),
            C.byref(ll) if want_ll else None)
        assert rc == 0
        out["rank"] = rank.value
        if want_ll:
            out["ll"] = ll.value
        return out


class RsOp(C.Structure):
    _fields_ = [("op", C.c_int32), ("reserved", C.c_int32), ("p0", C.c_double), ("p1", C.c_double)]


def _prog(ops, params):
    arr = (RsOp * len(ops))()
    for k, op in enumerate(ops):
        arr[k].op = int(op)
        arr[k].reserved = 0
        arr[k].p0 = float(params[2 * k])
        arr[k].p1 = float(params[2 * k + 1])
    return arr, C.c_int(len(ops))


def group_keys(feats, grouper_kind, grouper_arg=0.0):
    """Integer group key per feature for the shim groupers (oracle/ref_shim/ref_common.h)."""
    x0 = _feats(feats)[:, 0]
    if grouper_kind == 0:
        return np.arange(len(x0), dtype=np.int64)
    if grouper_kind == 1:
        return (np.trunc(x0).astype(np.int64) % int(grouper_arg)).astype(np.int64)
    return np.floor(x0 * grouper_arg).astype(np.int64)


class Restate:
    """The plain-C restatement (oracle/restate.c); covariance given as a postfix (ops, params)."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(RESTATE_SO):
                build("restate")
            lib = C.CDLL(RESTATE_SO)
            lib.rs_cov_eval.restype = C.c_double
            lib.rs_ldlt_logdet.restype = C.c_double
            lib.rs_nll_dense.restype = C.c_double
            lib.rs_gp_nll.restype = C.c_double
            lib.rs_group_indexers.restype = C.c_int64
            lib.rs_partition_triangular.restype = C.c_int64
            lib.rs_indices_complement.restype = C.c_int64
            cls._lib = lib
        return cls._lib

    @classmethod
    def cov_eval(cls, ops, params, x, y):
        prog, nops = _prog(ops, params)
        x = np.atleast_1d(np.asarray(x, dtype=np.float64))
        y = np.atleast_1d(np.asarray(y, dtype=np.float64))
        return cls.lib().rs_cov_eval(prog, nops, _d(x), _d(y), C.c_int(len(x)))

    @classmethod
    def gram_sym(cls, ops, params, feats):
        prog, nops = _prog(ops, params)
        x = _feats(feats)
        n, dim = x.shape
        out = np.empty((n, n), order="F")
        cls.lib().rs_gram_sym(prog, nops, _d(x), C.c_int64(n), C.c_int(dim), _d(out))
        return out

    @classmethod
    def gram_cross(cls, ops, params, fx, fy):
        prog, nops = _prog(ops, params)
        x, y = _feats(fx), _feats(fy)
        out = np.empty((x.shape[0], y.shape[0]), order="F")
        cls.lib().rs_gram_cross(prog, nops, _d(x), C.c_int64(x.shape[0]), _d(y),
                                C.c_int64(y.shape[0]), C.c_int(x.shape[1]), _d(out))
        return out

    @classmethod
    def gram_diag(cls, ops, params, feats):
        prog, nops = _prog(ops, params)
        x = _feats(feats)
        out = np.empty(x.shape[0])
        cls.lib().rs_gram_diag(prog, nops, _d(x), C.c_int64(x.shape[0]), C.c_int(x.shape[1]),
                               _d(out))
        return out

    @classmethod
    def ldlt(cls, A, rhs=None, want_inverse_diagonal=False):
        LD = np.array(A, dtype=np.float64, order="F", copy=True)
        n = LD.shape[0]
        tr = np.empty(n, dtype=np.int64)
        rc = cls.lib().rs_ldlt(_d(LD), C.c_int64(n), _i(tr))
        out = {"ldlt": LD, "transpositions": tr, "D": np.diag(LD).copy(), "rc": rc}
        out["logdet"] = cls.lib().rs_ldlt_logdet(_d(LD), C.c_int64(n))
        if rhs is not None:
            B = np.array(np.asarray(rhs, dtype=np.float64).reshape(n, -1), order="F", copy=True)
            k = B.shape[1]
            S = B.copy(order="F")
            cls.lib().rs_ldlt_solve(_d(LD), _i(tr), C.c_int64(n), _d(B), C.c_int64(k))
            cls.lib().rs_ldlt_sqrt_solve(_d(LD), _i(tr), C.c_int64(n), _d(S), C.c_int64(k))
            out["solve"], out["sqrt_solve"] = B, S
        if want_inverse_diagonal:
            inv = np.empty(n)
            cls.lib().rs_ldlt_inverse_diagonal(_d(LD), _i(tr), C.c_int64(n), _d(inv))
            out["inverse_diagonal"] = inv
        return out

    @classmethod
    def inverse_blocks(cls, A, groups):
        f = cls.ldlt(A)
        n = f["ldlt"].shape[0]
        indices = np.concatenate([np.asarray(g, dtype=np.int64) for g in groups])
        offsets = np.zeros(len(groups) + 1, dtype=np.int64)
        offsets[1:] = np.cumsum([len(g) for g in groups])
        out = np.empty(int(sum(len(g) ** 2 for g in groups)))
        cls.lib().rs_ldlt_inverse_blocks(_d(f["ldlt"]), _i(f["transpositions"]), C.c_int64(n),
                                         _i(indices), _i(offsets), C.c_int64(len(groups)), _d(out))
        blocks, c = [], 0
        for g in groups:
            k = len(g)
            blocks.append(out[c:c + k * k].reshape(k, k, order="F"))
            c += k * k
        return blocks

    @classmethod
    def nll_dense(cls, deviation, cov):
        d = np.ascontiguousarray(deviation, dtype=np.float64)
        c = np.asfortranarray(cov, dtype=np.float64)
        return cls.lib().rs_nll_dense(_d(d), _d(c), C.c_int64(len(d)))

    @classmethod
    def gp_fit(cls, ops, params, feats, y, yvar=None, want_factor=False):
        prog, nops = _prog(ops, params)
        x = _feats(feats)
        n, dim = x.shape
        y = np.ascontiguousarray(y, dtype=np.float64)
        yv = None if yvar is None else np.ascontiguousarray(yvar, dtype=np.float64)
        out = {"information": np.empty(n)}
        if want_factor:
            out["ldlt"] = np.empty((n, n), order="F")
            out["transpositions"] = np.empty(n, dtype=np.int64)
        cls.lib().rs_gp_fit(prog, nops, _d(x), C.c_int64(n), C.c_int(dim), _d(y), _d(yv),
                            _d(out["information"]), _d(out.get("ldlt")),
                            _i(out.get("transpositions")))
        return out

    @classmethod
    def gp_predict(cls, ops, params, feats, y, test, what, yvar=None):
        prog, nops = _prog(ops, params)
        x, t = _feats(feats), _feats(test)
        n, dim = x.shape
        pn = t.shape[0]
        y = np.ascontiguousarray(y, dtype=np.float64)
        yv = None if yvar is None else np.ascontiguousarray(yvar, dtype=np.float64)
        mean = np.empty(pn)
        var = np.empty(pn) if what == 1 else None
        cov = np.empty((pn, pn), order="F") if what == 2 else None
        cls.lib().rs_gp_predict(prog, nops, _d(x), C.c_int64(n), C.c_int(dim), _d(y), _d(yv),
                                _d(t), C.c_int64(pn), C.c_int(what), _d(mean), _d(var), _d(cov))
        return mean, var, cov

    @classmethod
    def gp_nll(cls, ops, params, feats, y):
        prog, nops = _prog(ops, params)
        x = _feats(feats)
        y = np.ascontiguousarray(y, dtype=np.float64)
        return cls.lib().rs_gp_nll(prog, nops, _d(x), C.c_int64(x.shape[0]), C.c_int(x.shape[1]),
                                   _d(y))

    @classmethod
    def group_indexers(cls, keys_per_item):
        gk = np.ascontiguousarray(keys_per_item, dtype=np.int64)
        n = len(gk)
        keys = np.empty(max(n, 1), dtype=np.int64)
        offsets = np.empty(n + 1, dtype=np.int64)
        indices = np.empty(max(n, 1), dtype=np.int64)
        g = cls.lib().rs_group_indexers(_i(gk), C.c_int64(n), _i(keys), _i(offsets), _i(indices))
        return keys[:g].copy(), offsets[:g + 1].copy(), indices[:n]

    @classmethod
    def partition_triangular(cls, n, count):
        out = np.empty(2 * count, dtype=np.int64)
        k = cls.lib().rs_partition_triangular(C.c_int64(n), C.c_int64(count), _i(out))
        return out[:2 * k].reshape(k, 2)

    @classmethod
    def indices_complement(cls, indices, n):
        idx = np.ascontiguousarray(indices, dtype=np.int64)
        out = np.empty(n, dtype=np.int64)
        k = cls.lib().rs_indices_complement(_i(idx), C.c_int64(len(idx)), C.c_int64(n), _i(out))
        return out[:k].copy()

    @classmethod
    def linspace(cls, a, b, n):
        out = np.empty(n)
        cls.lib().rs_linspace(C.c_double(a), C.c_double(b), C.c_int64(n), _d(out))
        return out

    @classmethod
    def gp_cv(cls, ops, params, feats, y, keys_per_item, what=1, want_score=False):
        prog, nops = _prog(ops, params)
        x = _feats(feats)
        n, dim = x.shape
        y = np.ascontiguousarray(y, dtype=np.float64)
        _, offsets, indices = cls.group_indexers(keys_per_item)
        sizes = np.diff(offsets)
        mean = np.empty(n)
        var = np.empty(n) if what == 1 else None
        joint = np.empty(int((sizes ** 2).sum())) if what == 2 else None
        score = C.c_double()
        cls.lib().rs_gp_cv(prog, nops, _d(x), C.c_int64(n), C.c_int(dim), _d(y), _i(indices),
                           _i(offsets), C.c_int64(len(sizes)), C.c_int(what), _d(mean), _d(var),
                           _d(joint), C.byref(score) if want_score else None)
        return mean, var, joint, (score.value if want_score else None)

    @classmethod
    def sparse_gp(cls, ops, params, feats, y, inducing, keys_per_item, test=None, what=-1,
                  yvar=None, measurement_nugget=1e-8, inducing_nugget=1e-8, want_ll=False,
                  fu=None, uu=None):
        """fu / uu: optional (ops, params) of the K_fu and K_uu (= prediction) programs when they
        differ from the K_ff one (MeasurementOnly terms, sparse_gp.hpp:646-679)."""
        prog, nops = _prog(ops, params)
        prog_fu, nops_fu = _prog(*fu) if fu is not None else (prog, nops)
        prog_uu, nops_uu = _prog(*uu) if uu is not None else (prog, nops)
        x = np.ascontiguousarray(feats, dtype=np.float64).ravel()
        n = len(x)
        u = np.ascontiguousarray(inducing, dtype=np.float64)
        m = len(u)
        y = np.ascontiguousarray(y, dtype=np.float64)
        yv = None if yvar is None else np.ascontiguousarray(yvar, dtype=np.float64)
        t = None if test is None else np.ascontiguousarray(test, dtype=np.float64).ravel()
        pn = 0 if t is None else len(t)
        _, offsets, indices = cls.group_indexers(keys_per_item)
        out = {"information": np.empty(m)}
        if what >= 0:
            out["mean"] = np.empty(pn)
        if what == 1:
            out["var"] = np.empty(pn)
        if what == 2:
            out["cov"] = np.empty((pn, pn), order="F")
        ll = C.c_double()
        cls.lib().rs_sparse_gp2(prog, nops, prog_fu, nops_fu, prog_uu, nops_uu, _d(x), C.c_int64(n),
                                _d(y), _d(yv), _d(u), C.c_int64(m), _i(indices), _i(offsets),
                                C.c_int64(len(offsets) - 1), C.c_double(measurement_nugget),
                                C.c_double(inducing_nugget), _d(t), C.c_int64(pn), C.c_int(what),
                                _d(out["information"]), _d(out.get("mean")), _d(out.get("var")),
                                _d(out.get("cov")), C.byref(ll) if want_ll else None)
        if want_ll:
            out["ll"] = ll.value
        return out

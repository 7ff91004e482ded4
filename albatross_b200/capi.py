"""ctypes binding of ``libalbatross_b200.so`` (the C ABI in ``include/albatross_b200.h``).

This is plumbing for the Python-side harness (tests, bench.py, smoke); the product is the shared
library itself and the C++ trait layer in ``albatross_b200/include``.  There is no CPU fallback:
if the library is missing or no CUDA device is visible the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# ALBATROSS_B200_LIB: development aid (tools/sweep.sh loads differently-tuned builds of the library)
LIB_PATH = os.environ.get("ALBATROSS_B200_LIB") or os.path.join(HERE, "csrc", "libalbatross_b200.so")

# opcodes (include/albatross_b200.h)
SE, EXP, M32, M52, CONST, NOISE, SUM, PROD, POLY = 1, 2, 3, 4, 5, 6, 7, 8, 9


def bench_program(name):
    """(ops, params) of the postfix programs BASELINE.json's configs use (timing tools / bench):
    "se_noise" = SE(1, 1) + IndependentNoise(0.1) (configs[2..4]); "se_m52" = SE(2, 1.5) + Matern52(3, 0.7)
    (configs[1]); "se_m52_noise" = the latter + IndependentNoise(0.1)."""
    if name == "se_noise":
        return [SE, NOISE, SUM], [1.0, 1.0, 0.1, 0.0, 0.0, 0.0]
    if name == "se_m52":
        return [SE, M52, SUM], [2.0, 1.5, 3.0, 0.7, 0.0, 0.0]
    if name == "se_m52_noise":
        return [SE, M52, SUM, NOISE, SUM], [2.0, 1.5, 3.0, 0.7, 0.0, 0.0, 0.1, 0.0, 0.0, 0.0]
    raise ValueError(name)
MEAN, MARGINAL, JOINT = 0, 1, 2
GRAM_FULL, GRAM_LOWER_ONLY = 0, 1

STATUS = {0: "AB_OK", 1: "AB_ERR_INVALID", 2: "AB_ERR_CUDA", 3: "AB_ERR_ALLOC",
          4: "AB_ERR_NOT_PD", 5: "AB_ERR_NCCL", 6: "AB_ERR_UNSUPPORTED"}


class AbError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"{STATUS.get(status, status)}: {message}")
        self.status = status


class AbOp(C.Structure):
    _fields_ = [("op", C.c_int32), ("reserved", C.c_int32), ("p0", C.c_double), ("p1", C.c_double)]


class PhaseTimes(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("h2d_ms", "gram_ms", "factor_ms", "solve_ms",
                                           "reduce_ms", "predict_ms", "d2h_ms", "total_ms")]
    _fields_.append(("kernel_launches", C.c_int64))

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)
_lib = None

# every symbol include/albatross_b200.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = [
    "ab_create", "ab_create_on_stream", "ab_destroy", "ab_last_error", "ab_version",
    "ab_device_count", "ab_synchronize", "ab_timings", "ab_reset_counters", "ab_trim",
    "ab_matrix_upload", "ab_matrix_alloc", "ab_matrix_download", "ab_matrix_download_block",
    "ab_matrix_dims", "ab_matrix_free", "ab_matrix_device_ptr", "ab_matrix_add_diag",
    "ab_gram_sym", "ab_gram_cross", "ab_gram_diag", "ab_gram_sym_d", "ab_gram_cross_d",
    "ab_potrf", "ab_factor_free", "ab_factor_rows", "ab_factor_info", "ab_factor_solve",
    "ab_factor_sqrt_solve", "ab_factor_logdet", "ab_factor_nll", "ab_factor_inverse_diagonal",
    "ab_factor_inverse_blocks", "ab_factor_export_packed", "ab_gp_fit", "ab_gp_nll",
    "ab_gp_fit_nll", "ab_gp_predict", "ab_gp_cv", "ab_gp_fit_d", "ab_gp_nll_d",
    "ab_group_indexers", "ab_gemm", "ab_gp_cv_shard", "ab_gp_cv_scores", "ab_gp_predict2",
    "ab_sparse_fit", "ab_sparse_free", "ab_sparse_info", "ab_sparse_log_likelihood",
    "ab_sparse_predict", "ab_sparse_export_R", "ab_sparse_fit2", "ab_sparse_log_likelihood2",
    "ab_sparse_predict2", "ab_factor_sqrt_product", "ab_factor_sqrt_transpose_solve",
    "ab_factor_sqrt_transpose", "ab_factor_diagonal_sqrt", "ab_qr_r", "ab_gp_update", "ab_factor_import_packed",
    "ab_dist_unique_id", "ab_dist_init", "ab_dist_finalize", "ab_dist_info", "ab_dist_gp_fit",
    "ab_dist_factor_free", "ab_dist_fit_breakdown", "ab_dist_factor_broadcast", "ab_dist_block_owner", "ab_dist_gram_rows", "ab_dist_gp_cv",
    "ab_partition_triangular",
]
DIST_ID_BYTES = 128


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (albatross_b200 has no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _lib.ab_last_error.restype = C.c_char_p
    return _lib


def _check(status):
    if status != 0:
        raise AbError(status, lib().ab_last_error().decode(errors="replace"))


def program(ops, params):
    """Postfix covariance program from opcode list + flat (p0, p1) per op."""
    arr = (AbOp * len(ops))()
    for k, op in enumerate(ops):
        arr[k].op = int(op)
        arr[k].reserved = 0
        arr[k].p0 = float(params[2 * k])
        arr[k].p1 = float(params[2 * k + 1])
    return arr, C.c_int(len(ops))


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _feats(x):
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    if x.ndim == 1:
        x = x.reshape(-1, 1)
    return x


def _vec(v):
    return None if v is None else np.ascontiguousarray(v, dtype=np.float64)


def group_indexers(item_keys):
    """Host-side integer contract (group_by(...).indexers()): returns keys, offsets, indices."""
    gk = np.ascontiguousarray(item_keys, dtype=np.int64)
    n = len(gk)
    keys = np.empty(max(n, 1), dtype=np.int64)
    offsets = np.empty(n + 1, dtype=np.int64)
    indices = np.empty(max(n, 1), dtype=np.int64)
    g = C.c_int64()
    _check(lib().ab_group_indexers(_i(gk), C.c_int64(n), _i(keys), _i(offsets), _i(indices),
                                   C.byref(g)))
    return keys[:g.value].copy(), offsets[:g.value + 1].copy(), indices[:n].copy()


def partition_triangular(n, count):
    """detail::partition_triangular (indexing/block.hpp:25-44): (count, 2) array of [start, end)."""
    out = np.empty(2 * count, dtype=np.int64)
    _check(lib().ab_partition_triangular(C.c_int64(n), C.c_int64(count), _i(out)))
    return out.reshape(count, 2)


class Matrix:
    """Device-resident column-major matrix (owning)."""

    def __init__(self, handle, ptr):
        self.h, self.ptr = handle, ptr

    @property
    def shape(self):
        r, c = C.c_int64(), C.c_int64()
        _check(lib().ab_matrix_dims(self.ptr, C.byref(r), C.byref(c)))
        return r.value, c.value

    def download(self):
        r, c = self.shape
        out = np.empty((r, c), order="F")
        _check(lib().ab_matrix_download(self.h.ptr, self.ptr, _d(out)))
        return out

    def download_block(self, row0, col0, rows, cols):
        out = np.empty((rows, cols), order="F")
        _check(lib().ab_matrix_download_block(self.h.ptr, self.ptr, C.c_int64(row0),
                                              C.c_int64(col0), C.c_int64(rows), C.c_int64(cols),
                                              _d(out)))
        return out

    def add_diag(self, d):
        d = _vec(d)
        _check(lib().ab_matrix_add_diag(self.h.ptr, self.ptr, _d(d)))

    def free(self):
        if self.ptr is not None:
            lib().ab_matrix_free(self.h.ptr, self.ptr)
            self.ptr = None

    def release(self):
        """Hands ownership to the caller (e.g. ab_potrf consumes the matrix)."""
        p, self.ptr = self.ptr, None
        return p

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Factor:
    """Device-resident Cholesky factor: the CovarianceRepresentation of the GP fit."""

    def __init__(self, handle, ptr):
        self.h, self.ptr = handle, ptr

    @property
    def n(self):
        n = C.c_int64()
        _check(lib().ab_factor_rows(self.ptr, C.byref(n)))
        return n.value

    def info(self):
        b = C.c_int64()
        _check(lib().ab_factor_info(self.ptr, C.byref(b)))
        return b.value

    def is_positive_definite(self):
        return self.info() < 0

    def _solve(self, fn, rhs):
        n = self.n
        rhs_f = np.asfortranarray(np.asarray(rhs, dtype=np.float64).reshape(n, -1))
        out = np.empty_like(rhs_f, order="F")
        _check(fn(self.h.ptr, self.ptr, _d(rhs_f), C.c_int64(rhs_f.shape[1]), _d(out)))
        return out.reshape(np.shape(rhs)) if np.ndim(rhs) == 1 else out

    def solve(self, rhs):
        return self._solve(lib().ab_factor_solve, rhs)

    def sqrt_solve(self, rhs):
        return self._solve(lib().ab_factor_sqrt_solve, rhs)

    def sqrt_product(self, rhs):
        return self._solve(lib().ab_factor_sqrt_product, rhs)

    def sqrt_transpose_solve(self, rhs):
        return self._solve(lib().ab_factor_sqrt_transpose_solve, rhs)

    def sqrt_transpose(self):
        out = np.empty((self.n, self.n), order="F")
        _check(lib().ab_factor_sqrt_transpose(self.h.ptr, self.ptr, _d(out)))
        return out

    def diagonal_sqrt(self):
        out = np.empty(self.n)
        _check(lib().ab_factor_diagonal_sqrt(self.h.ptr, self.ptr, _d(out)))
        return out

    def log_determinant(self):
        out = C.c_double()
        _check(lib().ab_factor_logdet(self.h.ptr, self.ptr, C.byref(out)))
        return out.value

    def nll(self, deviation):
        d = _vec(deviation)
        out = C.c_double()
        _check(lib().ab_factor_nll(self.h.ptr, self.ptr, _d(d), C.byref(out)))
        return out.value

    def inverse_diagonal(self):
        out = np.empty(self.n)
        _check(lib().ab_factor_inverse_diagonal(self.h.ptr, self.ptr, _d(out)))
        return out

    def inverse_blocks(self, groups):
        indices = np.concatenate([np.asarray(g, dtype=np.int64) for g in groups])
        offsets = np.zeros(len(groups) + 1, dtype=np.int64)
        offsets[1:] = np.cumsum([len(g) for g in groups])
        out = np.empty(int(sum(len(g) ** 2 for g in groups)))
        _check(lib().ab_factor_inverse_blocks(self.h.ptr, self.ptr, _i(indices), _i(offsets),
                                              C.c_int64(len(groups)), _d(out)))
        blocks, c = [], 0
        for g in groups:
            k = len(g)
            blocks.append(out[c:c + k * k].reshape(k, k, order="F"))
            c += k * k
        return blocks

    def export_packed(self):
        n = self.n
        LD = np.empty((n, n), order="F")
        tr = np.empty(n, dtype=np.int64)
        _check(lib().ab_factor_export_packed(self.h.ptr, self.ptr, _d(LD), _i(tr)))
        return LD, tr

    def free(self):
        if self.ptr is not None:
            lib().ab_factor_free(self.h.ptr, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Handle:
    """One GPU, one stream.  ``stream`` may be a raw cudaStream_t (int) to share torch's stream."""

    def __init__(self, device=0, stream=None):
        self.ptr = C.c_void_p()
        if stream is None:
            _check(lib().ab_create(C.byref(self.ptr), C.c_int(device)))
        else:
            _check(lib().ab_create_on_stream(C.byref(self.ptr), C.c_int(device),
                                             C.c_void_p(int(stream))))

    def close(self):
        if self.ptr:
            lib().ab_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        _check(lib().ab_synchronize(self.ptr))

    def trim(self):
        _check(lib().ab_trim(self.ptr))

    def timings(self):
        t = PhaseTimes()
        _check(lib().ab_timings(self.ptr, C.byref(t)))
        return t.as_dict()

    def reset_counters(self):
        _check(lib().ab_reset_counters(self.ptr))

    # -- matrices -----------------------------------------------------------------------------
    def upload(self, a):
        a = np.asarray(a, dtype=np.float64)
        if a.ndim == 1:
            a = a.reshape(-1, 1)
        a = np.asfortranarray(a)
        out = C.c_void_p()
        _check(lib().ab_matrix_upload(self.ptr, _d(a), C.c_int64(a.shape[0]),
                                      C.c_int64(a.shape[1]), C.byref(out)))
        return Matrix(self, out)

    def alloc(self, rows, cols):
        out = C.c_void_p()
        _check(lib().ab_matrix_alloc(self.ptr, C.c_int64(rows), C.c_int64(cols), C.byref(out)))
        return Matrix(self, out)

    def gemm(self, A, B, Cm, alpha=1.0, beta=0.0, trans_a=False, trans_b=False, lower=False):
        flags = (1 if trans_a else 0) | (2 if trans_b else 0) | (4 if lower else 0)
        _check(lib().ab_gemm(self.ptr, C.c_uint32(flags), C.c_double(alpha), A.ptr, B.ptr,
                             C.c_double(beta), Cm.ptr))

    def upload_features(self, feats):
        """Features (n, dim) -> dim x n device matrix (AoS preserved)."""
        x = _feats(feats)
        return self.upload(x.T)

    # -- gram ---------------------------------------------------------------------------------
    def gram_sym(self, ops, params, feats, flags=GRAM_FULL):
        prog, nops = program(ops, params)
        x = _feats(feats)
        out = C.c_void_p()
        _check(lib().ab_gram_sym(self.ptr, prog, nops, _d(x), C.c_int64(x.shape[0]),
                                 C.c_int(x.shape[1]), C.c_uint32(flags), C.byref(out)))
        return Matrix(self, out)

    def gram_sym_d(self, ops, params, feats_dev, flags=GRAM_FULL):
        prog, nops = program(ops, params)
        out = C.c_void_p()
        _check(lib().ab_gram_sym_d(self.ptr, prog, nops, feats_dev.ptr, C.c_uint32(flags),
                                   C.byref(out)))
        return Matrix(self, out)

    def gram_cross_d(self, ops, params, fx_dev, fy_dev):
        prog, nops = program(ops, params)
        out = C.c_void_p()
        _check(lib().ab_gram_cross_d(self.ptr, prog, nops, fx_dev.ptr, fy_dev.ptr, C.byref(out)))
        return Matrix(self, out)

    def gram_cross(self, ops, params, fx, fy):
        prog, nops = program(ops, params)
        x, y = _feats(fx), _feats(fy)
        out = C.c_void_p()
        _check(lib().ab_gram_cross(self.ptr, prog, nops, _d(x), C.c_int64(x.shape[0]), _d(y),
                                   C.c_int64(y.shape[0]), C.c_int(x.shape[1]), C.byref(out)))
        return Matrix(self, out)

    def gram_diag(self, ops, params, feats):
        prog, nops = program(ops, params)
        x = _feats(feats)
        out = np.empty(x.shape[0])
        _check(lib().ab_gram_diag(self.ptr, prog, nops, _d(x), C.c_int64(x.shape[0]),
                                  C.c_int(x.shape[1]), _d(out)))
        return out

    # -- factor -------------------------------------------------------------------------------
    def potrf(self, matrix, allow_not_pd=False):
        out = C.c_void_p()
        status = lib().ab_potrf(self.ptr, matrix.release(), C.byref(out))
        f = Factor(self, out) if out else None
        if status == 4 and allow_not_pd:
            return f
        _check(status)
        return f

    # -- exact GP -----------------------------------------------------------------------------
    def gp_fit(self, ops, params, feats, y, yvar=None, want_information=True):
        prog, nops = program(ops, params)
        x = _feats(feats)
        y = _vec(y)
        yv = _vec(yvar)
        info = np.empty(x.shape[0]) if want_information else None
        out = C.c_void_p()
        _check(lib().ab_gp_fit(self.ptr, prog, nops, _d(x), C.c_int64(x.shape[0]),
                               C.c_int(x.shape[1]), _d(y), _d(yv), C.byref(out), _d(info)))
        return Factor(self, out), info

    def gp_nll(self, ops, params, feats, y):
        prog, nops = program(ops, params)
        x = _feats(feats)
        y = _vec(y)
        out = C.c_double()
        _check(lib().ab_gp_nll(self.ptr, prog, nops, _d(x), C.c_int64(x.shape[0]),
                               C.c_int(x.shape[1]), _d(y), C.byref(out)))
        return out.value

    def gp_fit_nll(self, ops, params, feats, y):
        prog, nops = program(ops, params)
        x = _feats(feats)
        y = _vec(y)
        info = np.empty(x.shape[0])
        out = C.c_void_p()
        nll = C.c_double()
        _check(lib().ab_gp_fit_nll(self.ptr, prog, nops, _d(x), C.c_int64(x.shape[0]),
                                   C.c_int(x.shape[1]), _d(y), C.byref(out), _d(info),
                                   C.byref(nll)))
        return Factor(self, out), info, nll.value

    def gp_fit_d(self, ops, params, feats_dev, y_dev, yvar_dev=None, want_information=True):
        prog, nops = program(ops, params)
        out = C.c_void_p()
        info = C.c_void_p()
        _check(lib().ab_gp_fit_d(self.ptr, prog, nops, feats_dev.ptr, y_dev.ptr,
                                 None if yvar_dev is None else yvar_dev.ptr, C.byref(out),
                                 C.byref(info) if want_information else None))
        return Factor(self, out), (Matrix(self, info) if want_information else None)

    def gp_update(self, factor, ops, params, train_feats, information_old, new_feats, y_new, yvar_new=None):
        """Returns (Factor of size n + p, information[n + p])."""
        prog, nops = program(ops, params)
        x, xn = _feats(train_feats), _feats(new_feats)
        io, y, yv = _vec(information_old), _vec(y_new), _vec(yvar_new)
        info = np.empty(x.shape[0] + xn.shape[0])
        out = C.c_void_p()
        _check(lib().ab_gp_update(self.ptr, factor.ptr, prog, nops, _d(x), C.c_int64(x.shape[0]),
                                  C.c_int(x.shape[1]), _d(io), _d(xn), C.c_int64(xn.shape[0]), _d(y), _d(yv),
                                  C.byref(out), _d(info)))
        return Factor(self, out), info

    def gp_nll_d(self, ops, params, feats_dev, y_dev):
        prog, nops = program(ops, params)
        out = C.c_double()
        _check(lib().ab_gp_nll_d(self.ptr, prog, nops, feats_dev.ptr, y_dev.ptr, C.byref(out)))
        return out.value

    def gp_predict(self, factor, ops, params, train_feats, information, test_feats, what):
        prog, nops = program(ops, params)
        x, t = _feats(train_feats), _feats(test_feats)
        info = _vec(information)
        p = t.shape[0]
        mean = np.empty(p)
        var = np.empty(p) if what == MARGINAL else None
        cov = np.empty((p, p), order="F") if what == JOINT else None
        _check(lib().ab_gp_predict(self.ptr, factor.ptr, prog, nops, _d(x), C.c_int64(x.shape[0]),
                                   C.c_int(x.shape[1]), _d(info), _d(t), C.c_int64(p),
                                   C.c_int(what), _d(mean), _d(var), _d(cov)))
        return mean, var, cov

    def gp_cv(self, factor, y, information, offsets, indices, what, want_score=False):
        y, info = _vec(y), _vec(information)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        indices = np.ascontiguousarray(indices, dtype=np.int64)
        n = len(y)
        sizes = np.diff(offsets)
        mean = np.empty(n)
        var = np.empty(n) if what == MARGINAL else None
        joint = np.empty(int((sizes ** 2).sum())) if what == JOINT else None
        score = C.c_double()
        _check(lib().ab_gp_cv(self.ptr, factor.ptr, _d(y), _d(info), _i(indices), _i(offsets),
                              C.c_int64(len(sizes)), C.c_int(what), _d(mean), _d(var), _d(joint),
                              C.byref(score) if want_score else None))
        return mean, var, joint, (score.value if want_score else None)

    def gp_cv_shard(self, factor, y, information, offsets, indices, what, shard, nshards,
                    want_score=False):
        y, info = _vec(y), _vec(information)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        indices = np.ascontiguousarray(indices, dtype=np.int64)
        n = len(y)
        mean = np.empty(n)
        var = np.empty(n) if what == MARGINAL else None
        score = C.c_double()
        _check(lib().ab_gp_cv_shard(self.ptr, factor.ptr, _d(y), _d(info), _i(indices),
                                    _i(offsets), C.c_int64(len(offsets) - 1), C.c_int(what),
                                    C.c_int(shard), C.c_int(nshards), _d(mean), _d(var),
                                    C.byref(score) if want_score else None))
        return mean, var, (score.value if want_score else None)

    # -- sparse GP ----------------------------------------------------------------------------
    def _sparse_args(self, ops, params, feats, y, yvar, inducing, offsets, indices,
                     measurement_nugget, inducing_nugget, fu=None, uu=None):
        """fu / uu: optional (ops, params) of the K_fu and K_uu programs when they differ from the
        K_ff one (MeasurementOnly terms): the ab_sparse_*2 entry points are used then."""
        progs = [program(ops, params)]
        if fu is not None or uu is not None:
            progs.append(program(*(fu if fu is not None else (ops, params))))
            progs.append(program(*(uu if uu is not None else (ops, params))))
        x, u = _feats(feats), _feats(inducing)
        y, yv = _vec(y), _vec(yvar)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        indices = np.ascontiguousarray(indices, dtype=np.int64)
        keep = (progs, x, u, y, yv, offsets, indices)
        flat = tuple(v for pr in progs for v in pr)
        args = (self.ptr,) + flat + (_d(x), C.c_int64(x.shape[0]), C.c_int(x.shape[1]), _d(y),
                                     _d(yv), _d(u), C.c_int64(u.shape[0]), _i(indices), _i(offsets),
                                     C.c_int64(len(offsets) - 1), C.c_double(measurement_nugget),
                                     C.c_double(inducing_nugget))
        return args, keep, u.shape[0], len(progs) == 3

    def sparse_fit(self, ops, params, feats, y, inducing, offsets, indices, yvar=None,
                   measurement_nugget=1e-8, inducing_nugget=1e-8, fu=None, uu=None):
        """Returns (SparseFit, information[m], log_likelihood)."""
        args, keep, m, three = self._sparse_args(ops, params, feats, y, yvar, inducing, offsets,
                                                 indices, measurement_nugget, inducing_nugget, fu, uu)
        out = C.c_void_p()
        info = np.empty(m)
        ll = C.c_double()
        fn = lib().ab_sparse_fit2 if three else lib().ab_sparse_fit
        _check(fn(*args, C.byref(out), _d(info), C.byref(ll)))
        del keep
        return SparseFit(self, out), info, ll.value

    def sparse_log_likelihood(self, ops, params, feats, y, inducing, offsets, indices, yvar=None,
                              measurement_nugget=1e-8, inducing_nugget=1e-8, fu=None, uu=None):
        args, keep, _, three = self._sparse_args(ops, params, feats, y, yvar, inducing, offsets,
                                                 indices, measurement_nugget, inducing_nugget, fu, uu)
        ll = C.c_double()
        fn = lib().ab_sparse_log_likelihood2 if three else lib().ab_sparse_log_likelihood
        _check(fn(*args, C.byref(ll)))
        del keep
        return ll.value

    # -- distributed group --------------------------------------------------------------------
    def dist_init(self, rank, world, unique_id):
        buf = (C.c_char * DIST_ID_BYTES).from_buffer_copy(bytes(unique_id))
        _check(lib().ab_dist_init(self.ptr, C.c_int(rank), C.c_int(world), buf))

    def dist_init_from_torch(self):
        """Bootstraps the library's own NCCL communicator through an initialised torch.distributed
        process group (albatross_b200/dist.py)."""
        from . import dist

        return dist.bootstrap(self)

    def dist_finalize(self):
        _check(lib().ab_dist_finalize(self.ptr))

    def dist_info(self):
        r, w = C.c_int(), C.c_int()
        _check(lib().ab_dist_info(self.ptr, C.byref(r), C.byref(w)))
        return r.value, w.value

    def dist_gp_fit(self, ops, params, feats, y, yvar=None, nb=0, want_information=True):
        """All ranks call with identical arguments.  Returns (DistFactor, information, nll)."""
        prog, nops = program(ops, params)
        x = _feats(feats)
        y, yv = _vec(y), _vec(yvar)
        info = np.empty(x.shape[0]) if want_information else None
        out = C.c_void_p()
        nll = C.c_double()
        _check(lib().ab_dist_gp_fit(self.ptr, prog, nops, _d(x), C.c_int64(x.shape[0]),
                                    C.c_int(x.shape[1]), _d(y), _d(yv), C.c_int64(nb),
                                    C.byref(out), _d(info), C.byref(nll)))
        return DistFactor(self, out), info, nll.value

    def import_packed(self, LD, transpositions=None):
        """Factor from Eigen::SerializableLDLT's packed form (lower triangle of LD + transpositions)."""
        LDf = np.asfortranarray(LD, dtype=np.float64)
        n = LDf.shape[0]
        tr = None if transpositions is None else np.ascontiguousarray(transpositions, dtype=np.int64)
        out = C.c_void_p()
        _check(lib().ab_factor_import_packed(self.ptr, _d(LDf), _i(tr), C.c_int64(n), C.byref(out)))
        return Factor(self, out)

    def qr_r(self, B):
        """R (upper triangular, P = I) of a thin QR of the host matrix B."""
        Bf = np.asfortranarray(B, dtype=np.float64)
        rows, cols = Bf.shape
        R = np.empty((cols, cols), order="F")
        _check(lib().ab_qr_r(self.ptr, _d(Bf), C.c_int64(rows), C.c_int64(cols), _d(R)))
        return R

    def dist_factor_broadcast(self, factor, root=0):
        """Rank `root` passes its Factor, the others None; every rank returns a Factor of the same L."""
        ptr = C.c_void_p(factor.ptr.value if factor is not None else None)
        _check(lib().ab_dist_factor_broadcast(self.ptr, C.byref(ptr), C.c_int(root)))
        return factor if factor is not None else Factor(self, ptr)

    def dist_fit_breakdown(self):
        """(wait_ms, panel_ms, steps) of the most recent dist_gp_fit on this rank."""
        w, p, n = C.c_double(), C.c_double(), C.c_int64()
        _check(lib().ab_dist_fit_breakdown(self.ptr, C.byref(w), C.byref(p), C.byref(n)))
        return w.value, p.value, n.value

    def dist_gram_rows(self, ops, params, feats):
        """This rank's row block of the symmetric Gram: returns (row0, Matrix[rows x n])."""
        prog, nops = program(ops, params)
        x = _feats(feats)
        r0, nr = C.c_int64(), C.c_int64()
        out = C.c_void_p()
        _check(lib().ab_dist_gram_rows(self.ptr, prog, nops, _d(x), C.c_int64(x.shape[0]),
                                       C.c_int(x.shape[1]), C.byref(r0), C.byref(nr),
                                       C.byref(out)))
        return r0.value, Matrix(self, out)

    def dist_gp_cv(self, factor, y, information, offsets, indices, what, want_score=False):
        y, info = _vec(y), _vec(information)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        indices = np.ascontiguousarray(indices, dtype=np.int64)
        n = len(y)
        mean = np.empty(n)
        var = np.empty(n) if what == MARGINAL else None
        score = C.c_double()
        _check(lib().ab_dist_gp_cv(self.ptr, factor.ptr, _d(y), _d(info), _i(indices), _i(offsets),
                                   C.c_int64(len(offsets) - 1), C.c_int(what), _d(mean), _d(var),
                                   C.byref(score) if want_score else None))
        return mean, var, (score.value if want_score else None)


class SparseFit:
    """Device-resident Fit<SparseGPFit> (inducing features, K_uu factor, R factors, information)."""

    def __init__(self, handle, ptr):
        self.h, self.ptr = handle, ptr

    @property
    def m(self):
        m = C.c_int64()
        _check(lib().ab_sparse_info(self.ptr, C.byref(m), None))
        return m.value

    @property
    def log_likelihood(self):
        ll = C.c_double()
        _check(lib().ab_sparse_info(self.ptr, None, C.byref(ll)))
        return ll.value

    def predict(self, ops, params, test_feats, what, prior=None):
        """prior: optional (ops, params) of k(test, test) when it differs from the cross program."""
        prog, nops = program(ops, params)
        t = _feats(test_feats)
        p = t.shape[0]
        mean = np.empty(p)
        var = np.empty(p) if what == MARGINAL else None
        cov = np.empty((p, p), order="F") if what == JOINT else None
        if prior is None:
            _check(lib().ab_sparse_predict(self.h.ptr, self.ptr, prog, nops, _d(t), C.c_int64(p),
                                           C.c_int(what), _d(mean), _d(var), _d(cov)))
        else:
            pprog, pnops = program(*prior)
            _check(lib().ab_sparse_predict2(self.h.ptr, self.ptr, prog, nops, pprog, pnops, _d(t),
                                            C.c_int64(p), C.c_int(what), _d(mean), _d(var), _d(cov)))
        return mean, var, cov

    def export_R(self):
        m = self.m
        R = np.empty((m, m), order="F")
        _check(lib().ab_sparse_export_R(self.h.ptr, self.ptr, _d(R)))
        return R

    def free(self):
        if self.ptr:
            lib().ab_sparse_free(self.h.ptr, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DistFactor:
    """This rank's shard of a block-column-cyclic Cholesky factor."""

    def __init__(self, handle, ptr):
        self.h, self.ptr = handle, ptr

    def free(self):
        if self.ptr:
            lib().ab_dist_factor_free(self.h.ptr, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def dist_unique_id():
    buf = (C.c_char * DIST_ID_BYTES)()
    _check(lib().ab_dist_unique_id(buf))
    return bytes(buf.raw)


def dist_block_owner(block, world):
    r, lb = C.c_int(), C.c_int64()
    _check(lib().ab_dist_block_owner(C.c_int64(block), C.c_int(world), C.byref(r), C.byref(lb)))
    return r.value, lb.value


def device_count():
    n = C.c_int()
    status = lib().ab_device_count(C.byref(n))
    return n.value if status == 0 else 0

"""GPU parity of the DMMA GEMM family (ab_gemm) against numpy fp64: all four transpose forms,
ragged edges (m, n, k not multiples of the 128x128x16 tile), odd k on k-contiguous operands,
alpha/beta handling and the lower-triangle-only mode used by the trailing SYRK update."""
import numpy as np
import pytest

from tests.helpers import assert_close

pytestmark = pytest.mark.gpu

SHAPES = [(1, 1, 1), (7, 5, 3), (64, 64, 64), (128, 128, 16), (129, 127, 17), (130, 257, 33),
          (300, 200, 100), (65, 1, 1000), (1, 70, 513), (512, 384, 255)]


@pytest.mark.parametrize("ta", [False, True])
@pytest.mark.parametrize("tb", [False, True])
def test_gemm_forms(handle, ta, tb):
    rng = np.random.default_rng(0)
    for m, n, k in SHAPES:
        A = rng.standard_normal((k, m) if ta else (m, k))
        B = rng.standard_normal((n, k) if tb else (k, n))
        C0 = rng.standard_normal((m, n))
        want = -1.5 * (A.T if ta else A) @ (B.T if tb else B) + 0.5 * C0
        Cd = handle.upload(C0)
        handle.gemm(handle.upload(A), handle.upload(B), Cd, alpha=-1.5, beta=0.5, trans_a=ta,
                    trans_b=tb)
        assert_close(Cd.download(), want, 1e-13, f"gemm ta={ta} tb={tb} {m}x{n}x{k}")
        # beta == 0 must not read C (NaN-poisoned output buffer)
        Cn = handle.upload(np.full((m, n), np.nan))
        handle.gemm(handle.upload(A), handle.upload(B), Cn, alpha=1.0, beta=0.0, trans_a=ta,
                    trans_b=tb)
        assert_close(Cn.download(), (A.T if ta else A) @ (B.T if tb else B), 1e-13)


def test_gemm_lower_only(handle):
    rng = np.random.default_rng(1)
    for n, k in ((100, 40), (257, 130), (640, 64)):
        A = rng.standard_normal((n, k))
        C0 = rng.standard_normal((n, n))
        Cd = handle.upload(C0)
        handle.gemm(handle.upload(A), handle.upload(A), Cd, alpha=-1.0, beta=1.0, trans_b=True,
                    lower=True)
        got = Cd.download()
        want = C0 - A @ A.T
        assert_close(np.tril(got), np.tril(want), 1e-13, f"syrk {n}x{k}")
        # tiles strictly above the diagonal are untouched
        if n > 256:
            assert np.array_equal(got[:128, 256:], C0[:128, 256:])


def test_gemm_shape_errors(handle):
    from albatross_b200 import capi

    A, B, Cm = handle.alloc(4, 5), handle.alloc(6, 7), handle.alloc(4, 7)
    with pytest.raises(capi.AbError) as e:
        handle.gemm(A, B, Cm)
    assert e.value.status == 1

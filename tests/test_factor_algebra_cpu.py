"""The algebra `b200q_hinv_cholesky_upper` (csrc/linalg.cu) rests on, checked in float64 NumPy:

* the reference's factor (gptq.py:139-142: cholesky -> inverse -> upper cholesky of the inverse) is the
  unique upper-triangular U with positive diagonal and UᵀU = H⁻¹; with J the index reversal and
  JHJ = CᵀC (C upper), U = J·C⁻ᵀ·J — ONE factorization and ONE triangular inverse;
* the left-looking blocked Cholesky the device runs (block row j is brought up to date with one product
  over all finished rows, then its diagonal block is factored and its panel solved) is the Cholesky
  factor, for block sizes that do and do not divide K.
The device results themselves are compared with the oracle in tests/test_gptq_gpu.py and
tests/test_gptq_bench_shapes_gpu.py."""
import numpy as np
import pytest

from oracle import np_oracle as O
from tests.helpers import stable_seed


def _spd(k, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((4 * k, k)) * rng.uniform(0.2, 3.0, k)
    h = x.T @ x / (2 * k)
    return h + 0.01 * np.mean(np.diag(h)) * np.eye(k)


def _left_looking_upper_cholesky(a, nb):
    """C upper with CᵀC = a, block rows of ``nb``, the loop order of linalg.cu."""
    k = a.shape[0]
    c = np.triu(a).copy()
    for j0 in range(0, k, nb):
        j1 = min(j0 + nb, k)
        if j0:
            c[j0:j1, j0:] -= c[:j0, j0:j1].T @ c[:j0, j0:]          # one product over every finished row
        blk = np.triu(c[j0:j1, j0:j1])                                 # only the upper part of the block is valid
        d = np.linalg.cholesky(blk + np.triu(blk, 1).T).T
        c[j0:j1, j0:j1] = d
        if j1 < k:
            c[j0:j1, j1:] = np.linalg.solve(d.T, c[j0:j1, j1:])       # panel: C_jj^-T A[j, j+1:]
    return np.triu(c)


@pytest.mark.parametrize("k", [24, 40, 67])
def test_reversed_factorization_gives_the_references_upper_factor(k):
    h = _spd(k, stable_seed("factor_algebra", k))
    j = np.arange(k)[::-1]
    c = np.linalg.cholesky(h[np.ix_(j, j)]).T                        # JHJ = CᵀC
    u = np.linalg.inv(c).T[np.ix_(j, j)]                             # J C^-T J
    assert np.allclose(np.tril(u, -1), 0.0) and np.all(np.diag(u) > 0)
    assert np.allclose(u.T @ u, np.linalg.inv(h), rtol=1e-9, atol=1e-12)
    want = np.linalg.cholesky(np.linalg.inv(h)).T                    # the reference's route
    assert np.allclose(u, want, rtol=1e-8, atol=1e-12)
    ref32, ok = O.hinv_cholesky_upper(h.astype(np.float32), 0.0)     # the oracle's float32 restatement
    assert ok and np.allclose(ref32, u, rtol=2e-3, atol=2e-4 * np.abs(u).max())


@pytest.mark.parametrize("k,nb", [(40, 8), (67, 16), (128, 128), (50, 128)])
def test_left_looking_blocked_cholesky_is_the_cholesky_factor(k, nb):
    h = _spd(k, stable_seed("left_looking", k, nb))
    c = _left_looking_upper_cholesky(h, nb)
    assert np.allclose(c, np.linalg.cholesky(h).T, rtol=1e-10, atol=1e-12)


def _gptq_float64(w, u, block, left_looking):
    """GPTQ's block loop (gptq.py:153-208 with the transposed, non-zero triangle of U) in float64 with a
    fixed symmetric 4-bit grid per 16-row group — right-looking as the reference writes it (every
    finished block pushed into ALL later rows at once) or left-looking as csrc/gptq.cu runs it (a
    block's rows brought up to date right before the block, from the stored error rows)."""
    k, n = w.shape
    w = w.copy()
    codes = np.zeros((k, n), dtype=np.int64)
    err = np.zeros((k, n))
    scale = None
    for i1 in range(0, k, block):
        i2 = min(i1 + block, k)
        if left_looking and i1:
            w[i1:i2] -= u[:i1, i1:i2].T @ err[:i1]
        for i in range(i1, i2):
            if i % 16 == 0:
                scale = np.maximum(np.abs(w[i:i + 16]).max(axis=0), 1e-9) / 7.0
            q = np.clip(np.rint(w[i] / scale), -7, 7)
            codes[i] = q
            err[i] = (w[i] - q * scale) / u[i, i]
            w[i + 1:i2] -= np.outer(u[i, i + 1:i2], err[i])
        if not left_looking and i2 < k:
            w[i2:] -= u[i1:i2, i2:].T @ err[i1:i2]
    return codes, err


@pytest.mark.parametrize("k,block", [(64, 16), (96, 32), (80, 32)])
def test_left_looking_propagation_is_the_same_loop(k, block):
    rng = np.random.default_rng(stable_seed("left_looking_gptq", k, block))
    h = _spd(k, stable_seed("left_looking_gptq_h", k))
    u = np.linalg.cholesky(np.linalg.inv(h)).T
    w = rng.standard_normal((k, 24)) * 0.05
    c_right, e_right = _gptq_float64(w, u, block, left_looking=False)
    c_left, e_left = _gptq_float64(w, u, block, left_looking=True)
    assert np.array_equal(c_right, c_left)
    assert np.allclose(e_right, e_left, rtol=1e-9, atol=1e-12)

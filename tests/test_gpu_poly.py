"""GPU parity of the batch polynomial ops (C ABI) against the oracle's restatement of halo2's CPU routines."""
import numpy as np
import pytest

from tests import pyref
from tests.gpu_common import be  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 2, 255, 256, 257, 1000, 4096, 70001])
def test_batch_invert(be, orc, n):
    a = orc.fr_random_chacha(n, 900 + n)
    a[::7] = 0  # zeros must stay zero (ff::BatchInvert)
    assert np.array_equal(be.batch_invert(a), orc.batch_invert(a))


@pytest.mark.parametrize("n", [1, 2, 100, 4096, 4097, 100003])
def test_eval_polynomial(be, orc, n):
    poly = orc.fr_random_chacha(n, 30 + n)
    x = orc.fr_random_chacha(1, 31)[0]
    assert np.array_equal(be.eval_polynomial(poly, x), orc.eval_polynomial(poly, x))


@pytest.mark.parametrize("n", [2, 3, 256, 257, 258, 5000, 65537])
def test_kate_division(be, orc, n):
    a = orc.fr_random_chacha(n, 40 + n)
    b = orc.fr_random_chacha(1, 41)[0]
    assert np.array_equal(be.kate_division(a, b), orc.kate_division(a, b))


@pytest.mark.parametrize("n", [1, 2, 256, 257, 10000])
def test_grand_product(be, orc, n):
    a = orc.fr_random_chacha(n, 50 + n)
    ai = orc.fr_ints(a)
    want = [1]
    for v in ai[:-1]:
        want.append(want[-1] * v % pyref.R_MOD)
    assert orc.fr_ints(be.grand_product(a)) == want


def test_vec_ops(be, orc):
    n = 3001
    a = orc.fr_random_chacha(n, 60); b = orc.fr_random_chacha(n, 61); al = orc.fr_random_chacha(1, 62)[0]
    ai, bi, ali = orc.fr_ints(a), orc.fr_ints(b), orc.fr_ints(al)[0]
    assert orc.fr_ints(be.vec_mul(a, b)) == [x * y % pyref.R_MOD for x, y in zip(ai, bi)]
    assert orc.fr_ints(be.vec_axpy(a, al, b)) == [(x + ali * y) % pyref.R_MOD for x, y in zip(ai, bi)]
    assert orc.fr_ints(be.vec_scale(a, al)) == [x * ali % pyref.R_MOD for x in ai]

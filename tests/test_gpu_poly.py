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


@pytest.mark.parametrize("n", [2, 3, 64, 65, 66, 129, 256, 257, 258, 5000, 65537, 32769, 32770, (1 << 20) + 3])
def test_kate_division(be, orc, n):
    a = orc.fr_random_chacha(n, 40 + n)
    b = orc.fr_random_chacha(1, 41)[0]
    assert np.array_equal(be.kate_division(a, b), orc.kate_division(a, b))


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 256, 257, 10000, 65536, 65537, 100003])
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


def test_device_resident_variants_and_lincomb(be, orc):
    """_dev entry points on torch device buffers; lincomb = fold with powers of y (Horner over the polynomials)."""
    import torch
    n = 5000
    polys = [orc.fr_random_chacha(n, 80 + i) for i in range(4)]
    y = orc.fr_random_chacha(1, 90)[0]
    dev = torch.device("cuda", 0)
    d = [torch.from_numpy(p.view(np.int64)).to(dev) for p in polys]
    out = torch.empty((n, 4), dtype=torch.int64, device=dev)
    be.lincomb_dev([t.data_ptr() for t in d], y, out.data_ptr(), n)
    pi = [orc.fr_ints(p) for p in polys]; yi = orc.fr_ints(y)[0]
    want = [sum(pi[k][i] * pow(yi, k, pyref.R_MOD) for k in range(4)) % pyref.R_MOD for i in range(n)]
    assert orc.fr_ints(out.cpu().numpy().view(np.uint64)) == want
    x = orc.fr_random_chacha(1, 91)[0]
    assert np.array_equal(be.eval_polynomial_dev(d[0].data_ptr(), n, x), orc.eval_polynomial(polys[0], x))
    q = torch.empty((n - 1, 4), dtype=torch.int64, device=dev)
    be.kate_division_dev(d[0].data_ptr(), n, x, q.data_ptr())
    assert np.array_equal(q.cpu().numpy().view(np.uint64), orc.kate_division(polys[0], x))
    a = d[1].clone()
    be.batch_invert_dev(a.data_ptr(), n)
    assert np.array_equal(a.cpu().numpy().view(np.uint64), orc.batch_invert(polys[1]))
    z = torch.empty((n, 4), dtype=torch.int64, device=dev)
    be.grand_product_dev(d[2].data_ptr(), n, z.data_ptr())
    assert np.array_equal(z.cpu().numpy().view(np.uint64), be.grand_product(polys[2]))


@pytest.mark.parametrize("n", [1, 63, 64, 65, 5000, 70001])
def test_row_sharded_grand_product_pieces(be, orc, n):
    """spb_product_dev / spb_grand_product_seeded_dev: three row blocks scanned with exchanged totals give the same
    running product as one scan over all rows (the SURVEY 8e recipe, the collective replaced by a loop)."""
    import torch
    a = orc.fr_random_chacha(n, 0x5eed0900 + n)
    want = be.grand_product(a)
    bounds = [0, n // 3, (2 * n) // 3, n]
    da = torch.from_numpy(a.view(np.int64)).cuda()
    dz = torch.empty_like(da)
    R = pyref.R_MOD
    seed = 1
    for lo, hi in zip(bounds, bounds[1:]):
        blk = da[lo:hi]
        if hi > lo:
            be.grand_product_seeded_dev(blk.data_ptr(), hi - lo, orc.fr([seed])[0], dz[lo:hi].data_ptr())
        total = orc.fr_ints(be.product_dev(blk.data_ptr() if hi > lo else 0, hi - lo).reshape(1, 4))[0]
        prod = 1
        for v in orc.fr_ints(a[lo:hi]):
            prod = prod * v % R
        assert total == prod
        seed = seed * total % R
    assert np.array_equal(dz.cpu().numpy().view(np.uint64), want)


def test_fr_random_chacha_on_the_device_is_the_cpu_stream(be, orc):
    """spb_fr_random_chacha_dev draws `Fr::random(&mut ChaCha20Rng::from_seed(seed))` values in HBM: the same elements as the
    oracle's CPU keystream (which the seed-0 SRS secret pins against the verifier contracts' -tau G2), also across the 2^32
    word boundary of the block counter and for odd lengths; outputs are canonical residues."""
    import torch
    dev = torch.device("cuda", 0)
    for seed, first, n in ((0, 0, 1), (0x5eed, 0, 4097), (b"\x07" * 32, (1 << 32) - 5, 300), (2 ** 255 + 12345, (1 << 40) + 3, 777)):
        out = torch.empty((n, 4), dtype=torch.int64, device=dev)
        be.fr_random_chacha_dev(seed, first, out.data_ptr(), n)
        got = out.cpu().numpy().view(np.uint64)
        assert np.array_equal(got, orc.fr_random_chacha(n, seed, first)), (seed, first, n)
    assert np.array_equal(orc.fr_random_chacha(1, 0)[0], orc.srs_tau().reshape(4))   # draw 0 of seed 0 = the SRS secret

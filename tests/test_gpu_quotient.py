"""GPU parity of the quotient-numerator kernels (evaluate_h pieces) against the oracle's restatement, on synthetic
constraint systems. No reference-owned vector exists for this row (parity unpinned vs the Rust reference)."""
import random

import numpy as np
import pytest

from tests import pyref
from tests.gpu_common import be  # noqa: F401
from tests.quotient_common import random_program

pytestmark = pytest.mark.gpu


def _dev(torch, arr):
    return torch.from_numpy(np.ascontiguousarray(arr).view(np.int64)).cuda()


@pytest.mark.parametrize("seed,ncalc", [(1, 1), (2, 7), (3, 40), (4, 150)])
def test_graph_evaluator(be, orc, seed, ncalc):
    import torch
    rng = random.Random(seed)
    k, ek = 6, 8
    size, rot_scale = 1 << ek, 1 << (ek - k)
    n_fixed, n_advice, n_instance, n_chal, n_const = 3, 4, 1, 2, 5
    rotations = np.array([0, 1, -1, 2, 3, -(6 + 1)], dtype=np.int32)
    cols = lambda n, s: [orc.fr_random_chacha(size, 1000 * seed + s + i) for i in range(n)]
    fixed, advice, instance = cols(n_fixed, 10), cols(n_advice, 20), cols(n_instance, 30)
    constants = orc.fr_random_chacha(n_const, 40 + seed); constants[0] = 0; constants[1] = orc.fr([1])[0]
    challenges = orc.fr_random_chacha(n_chal, 50 + seed)
    bgty = orc.fr_random_chacha(4, 60 + seed)
    values0 = orc.fr_random_chacha(size, 70 + seed)
    prog = random_program(rng, ncalc, n_const, len(rotations), n_fixed, n_advice, n_instance, n_chal)
    want = orc.graph_evaluate(prog, ncalc, ncalc, constants, rotations, fixed, advice, instance, challenges, bgty, values0, rot_scale)
    df, da, di = [_dev(torch, a) for a in fixed], [_dev(torch, a) for a in advice], [_dev(torch, a) for a in instance]
    dv = _dev(torch, values0)
    be.graph_evaluate_dev(prog, ncalc, ncalc, constants, rotations, [t.data_ptr() for t in df], [t.data_ptr() for t in da], [t.data_ptr() for t in di],
                          challenges, bgty[0], bgty[1], bgty[2], bgty[3], dv.data_ptr(), size, rot_scale)
    assert np.array_equal(dv.cpu().numpy().view(np.uint64), want)


def test_graph_known_gate_matches_python(be, orc):
    """halo2-lib's basic gate q * (a + b*c - d) folded with y: checked against plain Python integers."""
    import torch
    from tests.quotient_common import ADD, SUB, MUL, HORNER, K_FIXED, K_ADVICE, K_INTER, K_PREV, K_Y
    size, rot_scale = 64, 4
    rotations = np.array([0, 1, 2, 3], dtype=np.int32)
    q = orc.fr_random_chacha(size, 1); a = orc.fr_random_chacha(size, 2)
    prog = np.array([MUL, 0, K_ADVICE, 0 | (1 << 16), K_ADVICE, 0 | (2 << 16),        # t0 = a[+1] * a[+2]
                     ADD, 1, K_ADVICE, 0, K_INTER, 0,                                  # t1 = a[0] + t0
                     SUB, 2, K_INTER, 1, K_ADVICE, 0 | (3 << 16),                      # t2 = t1 - a[+3]
                     MUL, 3, K_FIXED, 0, K_INTER, 2,                                   # t3 = q * t2
                     HORNER | (1 << 8), 4, K_PREV, 0, K_Y, 0, K_INTER, 3], dtype=np.uint32)  # prev * y + t3
    y = orc.fr_random_chacha(1, 3)[0]; prev = orc.fr_random_chacha(size, 4)
    dq, da, dv = _dev(torch, q), _dev(torch, a), _dev(torch, prev)
    zero = np.zeros((1, 4), dtype=np.uint64)
    be.graph_evaluate_dev(prog, 5, 5, zero, rotations, [dq.data_ptr()], [da.data_ptr()], [], zero, zero[0], zero[0], zero[0], y, dv.data_ptr(), size, rot_scale)
    qi, ai, pi, yi = orc.fr_ints(q), orc.fr_ints(a), orc.fr_ints(prev), orc.fr_ints(y)[0]
    R = pyref.R_MOD
    want = [(pi[i] * yi + qi[i] * (ai[i] + ai[(i + 4) % size] * ai[(i + 8) % size] - ai[(i + 12) % size])) % R for i in range(size)]
    assert orc.fr_ints(dv.cpu().numpy().view(np.uint64)) == want


@pytest.mark.parametrize("n_cols,chunk_len", [(1, 3), (5, 3), (6, 2), (13, 3)])
def test_permutation_constraints(be, orc, n_cols, chunk_len):
    import torch
    k, ek = 6, 8
    size, rot_scale = 1 << ek, 1 << (ek - k)
    n_sets = (n_cols + chunk_len - 1) // chunk_len
    mk = lambda s: orc.fr_random_chacha(size, s)
    z = [mk(100 + i) for i in range(n_sets)]; cv = [mk(200 + i) for i in range(n_cols)]; sg = [mk(300 + i) for i in range(n_cols)]
    l0, ll, la, values0 = mk(1), mk(2), mk(3), mk(4)
    beta, gamma, y = orc.fr_random_chacha(3, 5)
    wext = orc.fr([pyref.omega(ek)])[0]
    last_rotation = -(5 + 1)
    want = orc.permutation_constraints(values0, rot_scale, last_rotation, chunk_len, z, cv, sg, l0, ll, la, beta, gamma, y, wext)
    dz, dc, ds = [_dev(torch, a) for a in z], [_dev(torch, a) for a in cv], [_dev(torch, a) for a in sg]
    d0, dl, da, dv = _dev(torch, l0), _dev(torch, ll), _dev(torch, la), _dev(torch, values0)
    be.permutation_constraints_dev(dv.data_ptr(), size, rot_scale, last_rotation, chunk_len, [t.data_ptr() for t in dz], [t.data_ptr() for t in dc],
                                   [t.data_ptr() for t in ds], d0.data_ptr(), dl.data_ptr(), da.data_ptr(), beta, gamma, y, wext)
    assert np.array_equal(dv.cpu().numpy().view(np.uint64), want)


def test_lookup_constraints(be, orc):
    import torch
    size, rot_scale = 256, 4
    mk = lambda s: orc.fr_random_chacha(size, s)
    prod, pin, ptb, tv, l0, ll, la, values0 = [mk(i) for i in range(10, 18)]
    beta, gamma, y = orc.fr_random_chacha(3, 6)
    want = orc.lookup_constraints(values0, rot_scale, prod, pin, ptb, tv, l0, ll, la, beta, gamma, y)
    d = [_dev(torch, a) for a in (prod, pin, ptb, tv, l0, ll, la)]
    dv = _dev(torch, values0)
    be.lookup_constraints_dev(dv.data_ptr(), size, rot_scale, *[t.data_ptr() for t in d], beta, gamma, y)
    assert np.array_equal(dv.cpu().numpy().view(np.uint64), want)
    # first term checked against plain integers: v*y + (1 - z) * l0
    vi, zi, li, yi = orc.fr_ints(values0[:4]), orc.fr_ints(prod[:4]), orc.fr_ints(l0[:4]), orc.fr_ints(y)[0]
    first = orc.lookup_constraints(values0, rot_scale, prod, pin, ptb, tv, l0, np.zeros_like(ll), np.zeros_like(la), beta, gamma, y)
    assert first.shape == want.shape

"""CPU-only tests of the quotient-numerator and argument-prover term kernels: their per-row bodies
(spectre_b200/csrc/quotient.cuh, the code the CUDA kernels call) run serially in tests/hostemu -- with the emulated 32-bit
PTX limb arithmetic and with the 64-bit host path -- and are compared with the oracle's restatement of evaluate_h and of
the permutation / lookup grand products."""
import ctypes
import random

import numpy as np
import pytest

from tests import pyref
from tests.quotient_common import random_program
from tests.test_hostemu import _build, _p


@pytest.fixture(scope="module", params=["ptx", "native"])
def he(request):
    if request.param == "ptx":
        return _build("libhostemu_ptx.so", ["-DSPB_EMULATE_PTX"])
    return _build("libhostemu_native.so", [])


def _ptrs(arrs):
    arrs = [np.ascontiguousarray(a, dtype=np.uint64) for a in arrs]
    return (ctypes.c_void_p * max(1, len(arrs)))(*[a.ctypes.data for a in arrs]), arrs


@pytest.mark.parametrize("seed,ncalc,nslots", [(1, 1, 1), (2, 7, 3), (3, 40, 64), (4, 150, 7)])
def test_graph_evaluator_rows(he, orc, seed, ncalc, nslots):
    rng = random.Random(seed)
    k, ek = 5, 7
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
    got = values0.copy()
    pf, kf = _ptrs(fixed); pa, ka = _ptrs(advice); pi, ki = _ptrs(instance)
    scalars = np.ascontiguousarray(np.concatenate([bgty, challenges]))
    he.he_graph_evaluate(_p(prog), ctypes.c_uint32(ncalc), ctypes.c_uint32(ncalc), _p(constants), _p(rotations), pf, pa, pi, _p(scalars), _p(got),
                         ctypes.c_uint64(size), ctypes.c_int32(rot_scale), ctypes.c_uint32(nslots))
    assert np.array_equal(got, want)


@pytest.mark.parametrize("n_cols,chunk_len", [(1, 3), (5, 3), (6, 2), (13, 3)])
def test_permutation_constraint_rows(he, orc, n_cols, chunk_len):
    k, ek = 5, 7
    size, rot_scale = 1 << ek, 1 << (ek - k)
    n_sets = (n_cols + chunk_len - 1) // chunk_len
    mk = lambda s: orc.fr_random_chacha(size, s)
    z = [mk(100 + i) for i in range(n_sets)]; cv = [mk(200 + i) for i in range(n_cols)]; sg = [mk(300 + i) for i in range(n_cols)]
    l0, ll, la, values0 = mk(1), mk(2), mk(3), mk(4)
    beta, gamma, y = orc.fr_random_chacha(3, 5)
    wext = orc.fr([pyref.omega(ek)])[0]
    last_rotation = -(5 + 1)
    want = orc.permutation_constraints(values0, rot_scale, last_rotation, chunk_len, z, cv, sg, l0, ll, la, beta, gamma, y, wext)
    got = values0.copy()
    pz, kz = _ptrs(z); pc, kc = _ptrs(cv); ps, ks = _ptrs(sg)
    he.he_permutation_constraints(_p(got), ctypes.c_uint64(size), ctypes.c_int32(rot_scale), ctypes.c_int32(last_rotation), ctypes.c_uint32(n_sets),
                                  ctypes.c_uint32(chunk_len), pz, ctypes.c_uint32(n_cols), pc, ps, _p(l0), _p(ll), _p(la), _p(beta.copy()), _p(gamma.copy()),
                                  _p(y.copy()), _p(wext.copy()))
    assert np.array_equal(got, want)


def test_lookup_constraint_rows(he, orc):
    size, rot_scale = 128, 4
    mk = lambda s: orc.fr_random_chacha(size, s)
    prod, pin, ptb, tv, l0, ll, la, values0 = [mk(i) for i in range(10, 18)]
    beta, gamma, y = orc.fr_random_chacha(3, 6)
    want = orc.lookup_constraints(values0, rot_scale, prod, pin, ptb, tv, l0, ll, la, beta, gamma, y)
    got = values0.copy()
    he.he_lookup_constraints(_p(got), ctypes.c_uint64(size), ctypes.c_int32(rot_scale), _p(prod), _p(pin), _p(ptb), _p(tv), _p(l0), _p(ll), _p(la),
                             _p(beta.copy()), _p(gamma.copy()), _p(y.copy()))
    assert np.array_equal(got, want)


@pytest.mark.parametrize("k,n_cols,first_col", [(4, 1, 0), (9, 3, 2), (10, 2, 6)])
def test_grand_product_term_rows(he, orc, k, n_cols, first_col):
    """perm_terms_row with the kernel's block/thread split of omega^i, then batch inversion and the running product as the
    oracle's permutation_product does them; lookup_terms_row against lookup_product the same way"""
    n = 1 << k
    R = pyref.R_MOD
    values = [orc.fr_random_chacha(n, 100 + c) for c in range(n_cols)]
    sigma = [orc.fr_random_chacha(n, 200 + c) for c in range(n_cols)]
    beta, gamma = orc.fr_random_chacha(2, 300 + k)
    omega = orc.fr([pyref.omega(k)])[0]
    num = np.empty((n, 4), dtype=np.uint64); den = np.empty((n, 4), dtype=np.uint64)
    pv, kv = _ptrs(values); ps, ks = _ptrs(sigma)
    he.he_perm_terms(ctypes.c_uint32(k), pv, ps, ctypes.c_uint32(n_cols), ctypes.c_uint32(first_col), _p(beta.copy()), _p(gamma.copy()), _p(omega.copy()), _p(num), _p(den))
    one = orc.fr([1])[0]
    z_want, _ = orc.permutation_product(k, values, sigma, first_col, beta, gamma, np.zeros((0, 4), np.uint64), one)
    ni, di = orc.fr_ints(num), orc.fr_ints(orc.batch_invert(den))
    run, z = 1, []
    for a, b in zip(ni, di):
        z.append(run); run = run * a % R * b % R
    assert z == orc.fr_ints(z_want)
    arrs = [orc.fr_random_chacha(n, 500 + i) for i in range(4)]
    he.he_lookup_terms(*[_p(a) for a in arrs], _p(beta.copy()), _p(gamma.copy()), ctypes.c_uint64(n), _p(num), _p(den))
    z_want = orc.lookup_product(*arrs, beta, gamma, np.zeros((0, 4), np.uint64))
    ni, di = orc.fr_ints(num), orc.fr_ints(orc.batch_invert(den))
    run, z = 1, []
    for a, b in zip(ni, di):
        z.append(run); run = run * a % R * b % R
    assert z == orc.fr_ints(z_want)

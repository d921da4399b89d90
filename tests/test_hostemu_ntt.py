"""CPU-only test of the NTT kernel (spectre_b200/csrc/ntt.cuh): its barrier-separated phases are run serially by
tests/hostemu for the single-device plans and for the six-step plan across 2/4/8 emulated devices, and compared with
the oracle's best_fft / EvaluationDomain. Same plan and geometry code (ntt_make_plan, ntt_fill_pass) as the library."""
import ctypes

import numpy as np
import pytest

from tests import pyref
from tests.test_hostemu import _build, _p


@pytest.fixture(scope="module")
def he():
    return _build("libhostemu_native.so", [])   # 64-bit host arithmetic: the index/plan logic is what is under test


def he_ntt(he, a, k, omega, max_digit=11, tile_log=11, threads=256, g_log=0, n_in=0, n_out=0, pre3=None, post3=None, use_full=0):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    n = 1 << k
    out = np.zeros(((n_out or n), 4), dtype=np.uint64)
    rc = he.he_ntt(_p(a), _p(out), ctypes.c_uint32(k), _p(np.ascontiguousarray(omega, dtype=np.uint64)), ctypes.c_uint32(max_digit), ctypes.c_uint32(tile_log),
                   ctypes.c_uint32(threads), ctypes.c_uint32(g_log), ctypes.c_uint64(n_in), ctypes.c_uint64(n_out),
                   None if pre3 is None else _p(np.ascontiguousarray(pre3, dtype=np.uint64)), None if post3 is None else _p(np.ascontiguousarray(post3, dtype=np.uint64)),
                   ctypes.c_int(use_full))
    return out, rc


def _omega(orc, k):
    return orc.fr([pyref.omega(k)])[0]


@pytest.mark.parametrize("k,max_digit,tile_log,threads,full", [(0, 11, 11, 256, 0), (1, 11, 11, 256, 0), (5, 11, 10, 64, 0), (8, 11, 11, 256, 1), (9, 4, 6, 32, 0),
                                                                 (10, 5, 7, 64, 1), (12, 6, 8, 128, 0), (12, 11, 10, 256, 1), (13, 5, 6, 32, 0), (14, 7, 9, 512, 1)])
def test_single_device_plans(he, orc, k, max_digit, tile_log, threads, full):
    a = orc.fr_random_chacha(1 << k, 0x5eed0400 + k)
    w = _omega(orc, k)
    got, passes = he_ntt(he, a, k, w, max_digit, tile_log, threads, use_full=full)
    assert passes == max(1, -(-k // max_digit))
    assert np.array_equal(got, orc.best_fft(a, w, k))


@pytest.mark.parametrize("k,max_digit,tile_log,g_log", [(10, 5, 6, 1), (12, 6, 8, 2), (12, 4, 5, 1), (13, 5, 7, 3), (14, 7, 9, 2), (14, 5, 6, 3)])
def test_six_step_across_emulated_devices(he, orc, k, max_digit, tile_log, g_log):
    a = orc.fr_random_chacha(1 << k, 0x5eed0500 + k)
    w = _omega(orc, k)
    got, passes = he_ntt(he, a, k, w, max_digit, tile_log, 64, g_log=g_log, use_full=k % 2)
    assert passes >= 2
    assert np.array_equal(got, orc.best_fft(a, w, k))


@pytest.mark.parametrize("g_log", [0, 1, 2])
def test_fused_domain_variants(he, orc, g_log):
    """coeff_to_extended (zero padding + zeta pre-scale) and extended_to_coeff (post-scale + truncation) through the plans"""
    j, k = 4, 8
    d = orc.Domain(j, k)
    ek = d.extended_k
    coeff = orc.fr_random_chacha(1 << k, 11)
    one = orc.fr([1])[0]
    pre = np.stack([one, d.g_coset, d.g_coset_inv])
    ext, _ = he_ntt(he, coeff, ek, d.extended_omega, 4, 6, 32, g_log=g_log, n_in=1 << k, pre3=pre)
    assert np.array_equal(ext, d.coeff_to_extended(coeff))
    e = orc.fr_random_chacha(1 << ek, 12)
    I = orc.fr_ints
    div, zinv, z = I(d.extended_ifft_divisor)[0], I(d.g_coset_inv)[0], I(d.g_coset)[0]
    post = orc.fr([div, div * zinv % pyref.R_MOD, div * z % pyref.R_MOD])
    back, _ = he_ntt(he, e, ek, d.extended_omega_inv, 4, 6, 32, g_log=g_log, n_out=(1 << k) * (j - 1), post3=post)
    assert np.array_equal(back, d.extended_to_coeff(e))

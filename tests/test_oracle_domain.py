"""CPU tests of the oracle's best_fft / EvaluationDomain restatement against pure-Python big-int math
(no GPU; these also cross-check the coset ZETA convention the verifier KATs do not exercise)."""
import numpy as np
import pytest

from tests import pyref


@pytest.mark.parametrize("k", [1, 2, 3, 5, 7])
@pytest.mark.parametrize("threads", [1, 4, 8])
def test_best_fft_vs_python_dft(orc, k, threads):
    vals = [pow(5, 3 * i + 1, pyref.R_MOD) for i in range(1 << k)]
    w = pyref.omega(k)
    got = orc.fr_ints(orc.best_fft(orc.fr(vals), orc.fr([w])[0], k, threads=threads))
    assert got == pyref.dft(vals, w)


def test_best_fft_config1_size_threads_agree(orc):
    """BASELINE config 1 (Fr NTT 2^12 on CPU): iterative (threads >= n... ) and recursive paths agree."""
    k = 12
    a = orc.fr_random_chacha(1 << k, 0x5eed0001)
    w = orc.fr([pyref.omega(k)])[0]
    r1 = orc.best_fft(a, w, k, threads=1)
    r8 = orc.best_fft(a, w, k, threads=8)
    r4096 = orc.best_fft(a, w, k, threads=4096)  # log_n <= log_threads: the iterative branch
    assert np.array_equal(r1, r8) and np.array_equal(r1, r4096)
    # spot-check three outputs against the definition
    ai = orc.fr_ints(a)
    for i in (0, 1, 4095):
        assert orc.fr_ints(r1[i])[0] == sum(v * pow(pyref.omega(k), i * j, pyref.R_MOD) for j, v in enumerate(ai)) % pyref.R_MOD


def test_domain_constants_and_coset(orc):
    j, k = 4, 3
    d = orc.Domain(j, k)
    assert d.extended_k == 5  # 2^5 >= 8 * 3
    n = 1 << k
    wext = pyref.omega(d.extended_k)
    assert orc.fr_ints(d.extended_omega)[0] == wext
    assert orc.fr_ints(d.omega)[0] == pow(wext, 1 << (d.extended_k - k), pyref.R_MOD) == pyref.omega(k)
    assert orc.fr_ints(d.g_coset)[0] == pyref.ZETA
    assert orc.fr_ints(d.g_coset_inv)[0] == pow(pyref.ZETA, 2, pyref.R_MOD) == pow(pyref.ZETA, -1, pyref.R_MOD)
    assert orc.fr_ints(d.ifft_divisor)[0] == pow(n, -1, pyref.R_MOD)
    t = orc.fr_ints(d.t_evaluations)
    assert len(t) == 1 << (d.extended_k - k)
    for i, ti in enumerate(t):
        x = pyref.ZETA * pow(wext, i, pyref.R_MOD) % pyref.R_MOD
        assert ti == pow(pow(x, n, pyref.R_MOD) - 1, -1, pyref.R_MOD)
    coeffs = [pow(11, i + 1, pyref.R_MOD) for i in range(n)]
    ext = orc.fr_ints(d.coeff_to_extended(orc.fr(coeffs)))
    for i in range(1 << d.extended_k):
        x = pyref.ZETA * pow(wext, i, pyref.R_MOD) % pyref.R_MOD
        assert ext[i] == sum(c * pow(x, e, pyref.R_MOD) for e, c in enumerate(coeffs)) % pyref.R_MOD
    # lagrange_to_coeff inverts evaluation on the omega-domain
    evals = [sum(c * pow(pyref.omega(k), i * e, pyref.R_MOD) for e, c in enumerate(coeffs)) % pyref.R_MOD for i in range(n)]
    assert orc.fr_ints(d.lagrange_to_coeff(orc.fr(evals))) == coeffs
    # extended_to_coeff(coeff_to_extended(p)) == p padded to n*(j-1)
    back = orc.fr_ints(d.extended_to_coeff(orc.fr(ext)))
    assert back == coeffs + [0] * (n * (j - 1) - n)


def test_poly_helpers(orc):
    import random
    rng = random.Random(5)
    poly = [rng.randrange(pyref.R_MOD) for _ in range(33)]
    x = rng.randrange(pyref.R_MOD)
    assert orc.fr_ints(orc.eval_polynomial(orc.fr(poly), orc.fr([x])[0]))[0] == sum(c * pow(x, i, pyref.R_MOD) for i, c in enumerate(poly)) % pyref.R_MOD
    q = orc.fr_ints(orc.kate_division(orc.fr(poly), orc.fr([x])[0]))
    # (X - x) * q + p(x) == p
    px = sum(c * pow(x, i, pyref.R_MOD) for i, c in enumerate(poly)) % pyref.R_MOD
    rebuilt = [0] * len(poly)
    for i, qi in enumerate(q):
        rebuilt[i + 1] = (rebuilt[i + 1] + qi) % pyref.R_MOD
        rebuilt[i] = (rebuilt[i] - x * qi) % pyref.R_MOD
    rebuilt[0] = (rebuilt[0] + px) % pyref.R_MOD
    assert rebuilt == poly
    vals = [0, 1, 2, 0, pyref.R_MOD - 1] + [rng.randrange(pyref.R_MOD) for _ in range(20)]
    inv = orc.fr_ints(orc.batch_invert(orc.fr(vals)))
    assert inv == [pow(v, -1, pyref.R_MOD) if v else 0 for v in vals]

"""Pin the CPU oracle against every numeric artefact the reference commits for this path
(tests/golden/verifier_kats.json, extracted from contracts/snark-verifiers/*.sol by tools/extract_golden.py)."""
import numpy as np
import pytest

from tests import pyref


def h(x):
    return int(x, 16)


def test_primes_match_verifier(kats):
    assert h(kats["f_q_scalar_field_r"]["value"]) == pyref.R_MOD
    assert h(kats["f_p_base_field_p"]["value"]) == pyref.P_MOD


def test_generators(orc, kats):
    g = orc.affine_ints(orc.g1_generator())[0]
    assert list(g) == [h(v) for v in kats["g1_generator"]["xy"]]
    assert orc.g1_on_curve(orc.g1_generator())


def test_seed0_tau_and_s_g2(orc, kats):
    """ChaCha20Rng::from_seed([0;32]) -> Fr::random -> tau; tau*G2 == the verifier's pairing constant (y negated)."""
    tau = orc.fr_ints(orc.srs_tau())[0]
    assert tau == 0x1c59a59b6cff4308740943526ade1d8c09f71b337a67269cc89586bcdd6dfcba
    xc0, xc1, yc0, yc1 = orc.fq_ints(orc.srs_s_g2())
    for key in ("neg_s_g2_sync_step", "neg_s_g2_committee_update"):
        w = [h(v) for v in kats[key]["x_c1,x_c0,y_c1,y_c0"]]
        assert [xc1, xc0] == w[:2]
        assert [(-yc1) % pyref.P_MOD, (-yc0) % pyref.P_MOD] == w[2:]


@pytest.mark.slow
def test_range_table_commit_k23_via_best_multiexp(orc, kats):
    """commit_lagrange(range table) through the restated best_multiexp over the restated seed-0 g_lagrange
    equals the fixed-column commitment in the sync-step aggregation verifier (K=23, lookup_bits=19)."""
    kat = kats["range_table_commit_k23_bits19"]
    k, bits = kat["k"], kat["lookup_bits"]
    n_used = 1 << bits
    bases = orc.srs_g_lagrange(k, 0, n_used)
    coeffs = orc.fr(range(n_used))
    res = orc.affine_ints(orc.g1_to_affine(orc.best_multiexp(coeffs, bases)))[0]
    assert list(res) == [h(v) for v in kat["xy"]]
    # and the O(n) known-tau shortcut agrees
    short = orc.affine_ints(orc.commit_lagrange_known_tau(k, coeffs))[0]
    assert short == res


@pytest.mark.slow
def test_range_table_commit_k24_known_tau(orc, kats):
    """K=24, lookup_bits=23 (committee-update aggregation verifier) through the known-tau shortcut."""
    kat = kats["range_table_commit_k24_bits23"]
    k, bits = kat["k"], kat["lookup_bits"]
    vals = orc.fr_seq(1 << bits)
    res = orc.affine_ints(orc.commit_lagrange_known_tau(k, vals))[0]
    assert list(res) == [h(v) for v in kat["xy"]]

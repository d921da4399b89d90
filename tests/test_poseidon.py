"""Poseidon over Fr (spectre_b200/poseidon.py): the permutation against the known-answer vector published with the Poseidon
reference implementation, the parameter generation for the widths the reference uses, and the transcript's self-consistency
with the proof driver and the independent verifier. (The sponge / transcript framing is unpinned -- module docstring.)"""
import numpy as np
import pytest

from spectre_b200 import circuits, plonk, poseidon
from tests import plonk_verifier
from tests.plonk_oracle_engine import OracleEngine, SeededRng

R = poseidon.R_MOD


def test_permutation_known_answer_x5_254_3():
    """`poseidonperm_x5_254_3` of the reference implementation (t = 3, R_F = 8, R_P = 57 over BN254's scalar field -- the
    parameters of snark-verifier's PoseidonTranscript): permutation of (0, 1, 2)."""
    spec = poseidon.Spec(3, 8, 57)
    assert spec.permute([0, 1, 2]) == [
        0x115cc0f5e7d690413df64c6b9662e9cf2a3617f2743245519e19607a4417189a,
        0x0fca49b798923ab0239de1c9e7a4a9a2210312b6a2f616d18b5a87f9b628ae29,
        0x0e7ae82e40091e63cbd4f16a6d16310b3729d4b6e138fcf54110e2867045a30c,
    ]
    # first round constant and first MDS entry of that instance, as the reference script prints them
    assert spec.constants[0][0] == 0x0ee9a592ba9a9518d05986d656f40c2114c4993c11bb29938d21d47304cd8e6e
    assert spec.mds[0][0] == 0x109b7f411ba0e4c9b2b70caf5c36a7b194be7c11ad24378bfedb68592ba8118b


@pytest.mark.parametrize("t,r_p", [(3, 57), (12, 60)])
def test_parameters_are_a_valid_instance(t, r_p):
    """widths the reference uses: 3 (transcript) and 12 (committee commitment, lightclient-circuits/src/poseidon.rs:21-31: R_P =
    N_ROUNDS_PC[T - 2] = 60): canonical constants, an invertible Cauchy matrix, and a permutation that is a bijection on samples"""
    spec = poseidon.Spec(t, 8, r_p)
    assert len(spec.constants) == 8 + r_p and all(len(row) == t and all(0 <= c < R for c in row) for row in spec.constants)
    # Cauchy matrices are invertible: solve M x = e_0 by Gaussian elimination over Fr
    m = [row[:] + [1 if i == 0 else 0] for i, row in enumerate(spec.mds)]
    for col in range(t):
        piv = next(r for r in range(col, t) if m[r][col])
        m[col], m[piv] = m[piv], m[col]
        inv = pow(m[col][col], R - 2, R)
        m[col] = [v * inv % R for v in m[col]]
        for r in range(t):
            if r != col and m[r][col]:
                f = m[r][col]
                m[r] = [(a - f * b) % R for a, b in zip(m[r], m[col])]
    outs = {tuple(spec.permute([i] + [0] * (t - 1))) for i in range(8)}
    assert len(outs) == 8


def test_sponge_framing():
    """update buffers, squeeze absorbs RATE at a time; a short chunk and an exactly full buffer are framed differently, and
    splitting the same elements over several updates changes nothing"""
    a = poseidon.Poseidon(); a.update([1, 2, 3]); ha = a.squeeze()
    b = poseidon.Poseidon(); b.update([1]); b.update([2, 3]); assert b.squeeze() == ha
    c = poseidon.Poseidon(); c.update([1, 2]); hc = c.squeeze()
    d = poseidon.Poseidon(); d.update([1, 2, 1]); assert d.squeeze() != hc          # the padding 1 is not an element
    e = poseidon.Poseidon(); assert e.squeeze() != hc and 0 <= hc < R
    # one permutation by hand: state (2^64, 0, 0) + (5, padding 1)
    f = poseidon.Poseidon(); f.update([5])
    assert f.squeeze() == poseidon.Spec(3, 8, 57).permute([1 << 64, 5, 1])[1]


def test_point_compression_round_trip(orc):
    pts = orc.g1_fixed_base_mul(orc.fr_random_chacha(8, 31))
    for xy in orc.affine_ints(pts):
        b = poseidon.compress_g1(xy)
        assert len(b) == 32 and poseidon.decompress_g1(b) == tuple(xy)
    assert poseidon.decompress_g1(poseidon.compress_g1((0, 0))) == (0, 0)
    with pytest.raises(ValueError):
        poseidon.decompress_g1((4).to_bytes(32, "little"))                             # x = 4: x^3 + 3 = 67 is not a square mod p


def test_proof_over_the_poseidon_transcript_verifies(orc):
    """create_proof is transcript-agnostic: over PoseidonTranscriptWrite the multi-column shape proves, the independent verifier
    accepts it through PoseidonTranscriptRead, rejects a flipped byte, and the bytes differ from the Keccak-transcript proof
    (compressed points: shorter)."""
    from spectre_b200.transcript import EvmTranscriptWrite
    k, inst = 8, [4, 2]
    cs = circuits.halo2lib_shape(3, 2)
    fixed, adv, copies = circuits.halo2lib_witness(cs, k, inst, lookup_bits=4, groups=20, num_gate_advice=3, num_lookup_advice=2)
    E = OracleEngine(k, cs.degree())
    pk = plonk.keygen(E, cs, k, fixed, copies)
    proof = plonk.create_proof(E, pk, [inst], adv, SeededRng(9), poseidon.PoseidonTranscriptWrite(pk.vk_digest))
    evm = plonk.create_proof(E, pk, [inst], adv, SeededRng(9), EvmTranscriptWrite(pk.vk_digest))
    assert len(proof) < len(evm)
    tau = orc.fr_ints(orc.srs_tau().reshape(1, 4))[0]
    args = (cs, k, pk.vk_digest, pk.fixed_commitments, pk.sigma_commitments, [inst])
    assert plonk_verifier.verify(*args, proof, tau, transcript_read=poseidon.PoseidonTranscriptRead)
    bad = bytearray(proof); bad[len(bad) // 2] ^= 1
    with pytest.raises((AssertionError, ValueError)):
        plonk_verifier.verify(*args, bytes(bad), tau, transcript_read=poseidon.PoseidonTranscriptRead)

"""Synthetic circuits (constraint system + satisfying witness) for the proof driver: the workload of the parity tests and of
`bench.py --prove`.

`aggregation_shape()` is the constraint system of Spectre's aggregation circuits as their committed verifier contract
spells it out (contracts/snark-verifiers/sync_step_verifier.sol:507-590):
halo2-lib's single-advice-column layout with
  * one advice column queried at rotations 0..3 by the basic gate  q_gate * (a + b*c - d)            (.sol:507-510),
  * four fixed columns evaluated in the order: constants, range table, q_lookup, q_gate               (.sol evals 4-7),
  * one lookup  q_lookup * a  in  table                                                               (.sol:572-581),
  * one permutation over (constants, advice, instance)                                                (.sol:528-556),
so degree 5, four quotient pieces, six blinding factors, 19 evaluations -- the proof layout of the 0x720-byte calldata.
The witness below is synthetic (groups of four cells a + b*c = d chained by copy constraints); the real circuit's
fixed columns are unknown here, so the VK commitments differ from the contract's constants.
"""
import random

import numpy as np

from . import plonk
from .plonk import Advice, Fixed, Neg, Prod, Sum

R = plonk.R_MOD


def aggregation_shape():
    a = [Advice(0, r) for r in range(4)]
    gate = Prod(Sum(Sum(a[0], Prod(a[1], a[2])), Neg(a[3])), Fixed(3))
    return plonk.ConstraintSystem(num_fixed=4, num_advice=1, num_instance=1, gates=[gate],
                                  lookups=[([Prod(Advice(0), Fixed(2))], [Fixed(1)])],
                                  permutation=[("fixed", 0), ("advice", 0), ("instance", 0)],
                                  fixed_queries=[(0, 0), (1, 0), (2, 0), (3, 0)])


def aggregation_witness(cs, k, instances, lookup_bits, groups, seed=1, dense=False):
    """-> (fixed columns [4 x (n,4)], advice column (n,4), copies). instances: list of ints (instance column 0).
    dense: fill the unconstrained advice / constants rows (no gate, no lookup, no copy on them) with random residues, so
    the commitments see a full column as they do in the real circuit."""
    n = 1 << k
    usable = n - (cs.blinding_factors() + 1)
    groups = min(groups, (usable - 4) // 4)
    assert len(instances) <= groups and (1 << lookup_bits) <= usable
    rng = random.Random(seed)
    const, table, q_lookup, q_gate, adv = [0] * n, [0] * n, [0] * n, [0] * n, [0] * n
    for i in range(1 << lookup_bits):
        table[i] = i
    c0 = rng.randrange(R)
    const[0] = c0
    copies = [(("adv", 1), ("const", 0))]
    prev_d = None
    for g in range(groups):
        r = 4 * g
        a = rng.randrange(1 << lookup_bits)
        b = c0 if g == 0 else prev_d
        c = instances[g] if g < len(instances) else rng.randrange(R)
        d = (a + b * c) % R
        adv[r:r + 4] = [a, b, c, d]
        q_gate[r] = 1; q_lookup[r] = 1
        if g > 0:
            copies.append((("adv", r + 1), ("adv", r - 1)))
        if g < len(instances):
            copies.append((("adv", r + 2), ("inst", g)))
        prev_d = d
    col = {"const": 0, "adv": 1, "inst": 2}                 # index in cs.permutation
    copies = [((col[a[0]], a[1]), (col[b[0]], b[1])) for a, b in copies]

    def mont(vals):                                          # sparse-aware conversion: most rows are zero
        out = np.zeros((n, 4), dtype=np.uint64)
        for i, v in enumerate(vals):
            if v:
                out[i] = plonk.fr_mont(v)
        return out
    fixed, advice = [mont(const), mont(table), mont(q_lookup), mont(q_gate)], mont(adv)
    if dense:
        g = np.random.default_rng(seed)
        for col, lo in ((advice, 4 * groups), (fixed[0], 1)):
            fill = g.integers(0, 1 << 63, size=(usable - lo, 4), dtype=np.uint64)
            fill[:, 3] &= np.uint64((1 << 60) - 1)           # < 2^252 < r: valid Montgomery residues
            col[lo:usable] = fill
    return fixed, advice, copies


def wide_shape(num_advice=3):
    """A second shape: several advice columns (two permutation sets, inter-set terms), a two-column lookup compressed
    with theta, and a multiplication gate on column 1 -- exercises what the aggregation shape does not."""
    gates = []
    for c in range(num_advice):
        a = [Advice(c, r) for r in range(4)]
        gates.append(Prod(Sum(Sum(a[0], Prod(a[1], a[2])), Neg(a[3])), Fixed(c)))
    lookups = [([Prod(Advice(0), Fixed(num_advice)), Prod(Advice(1), Fixed(num_advice))], [Fixed(num_advice + 1), Fixed(num_advice + 2)])]
    perm = [("advice", c) for c in range(num_advice)] + [("fixed", num_advice + 3), ("instance", 0)]
    return plonk.ConstraintSystem(num_fixed=num_advice + 4, num_advice=num_advice, num_instance=1, gates=gates, lookups=lookups, permutation=perm)


def wide_witness(cs, k, instances, lookup_bits, groups, seed=2):
    n = 1 << k
    A = cs.num_advice
    usable = n - (cs.blinding_factors() + 1)
    groups = min(groups, (usable - 4) // 4)
    rng = random.Random(seed)
    fixed = [[0] * n for _ in range(cs.num_fixed)]
    adv = [[0] * n for _ in range(A)]
    # table: pairs (i, i*i + 1)
    for i in range(1 << lookup_bits):
        fixed[A + 1][i] = i; fixed[A + 2][i] = (i * i + 1) % R
    fixed[A + 3][0] = rng.randrange(R)
    copies = []
    for g in range(groups):
        r = 4 * g
        for c in range(A):
            a = rng.randrange(1 << lookup_bits)
            if c == 1:
                a = (adv[0][r] * adv[0][r] + 1) % R          # second lookup column: the table's second coordinate
            b = rng.randrange(R); cc = rng.randrange(R)
            if c == 0 and g < len(instances): cc = instances[g]
            if c == 2 and g == 0: b = fixed[A + 3][0]
            if c == 0 and g > 0: b = adv[A - 1][r - 1]
            adv[c][r:r + 4] = [a, b, cc, (a + b * cc) % R]
            fixed[c][r] = 1
        fixed[A][r] = 1                                       # q_lookup
        if g < len(instances): copies.append(((0, r + 2), (A + 1, g)))
        if g == 0 and A > 2: copies.append(((2, r + 1), (A, 0)))
        if g > 0: copies.append(((0, r + 1), (A - 1, r - 1)))

    def mont(vals):
        out = np.zeros((n, 4), dtype=np.uint64)
        for i, v in enumerate(vals):
            if v: out[i] = plonk.fr_mont(v)
        return out
    return [mont(f) for f in fixed], [mont(a) for a in adv], copies

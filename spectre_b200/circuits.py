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


# ---- halo2-lib's multi-column layout: the shape of Spectre's sync-step circuit -----------------------------------------
def halo2lib_shape(num_gate_advice=15, num_lookup_advice=2, spread=True):
    """The constraint-system shape SURVEY.md section 8 (table row 1) derives for the sync-step circuit from
    lightclient-circuits/config/sync_step_20.json:3-16 and the SHA spread config (sha256_flex/spread.rs:40-80,89):
    `num_gate_advice` basic-gate columns q_c * (a + b*c - d), `num_lookup_advice` dedicated range-lookup columns (no
    selector: degree-4 lookups), one two-column spread lookup compressed with theta, and one permutation over every advice
    column, the constants column and the instance column. With the defaults: 19 advice columns, 3 lookups, degree 4
    (extended domain 4n, 3 quotient pieces), 21 permutation columns in 11 sets of 2 -- the 45-MSM schedule of that row.
    advice columns: [gate 0..G) [lookup G..G+L) [spread dense, spread spread]; fixed: [q_gate 0..G) constants, range table,
    spread table (2 columns)."""
    G, L = num_gate_advice, num_lookup_advice
    A = G + L + (2 if spread else 0)
    gates = []
    for c in range(G):
        a = [Advice(c, r) for r in range(4)]
        gates.append(Prod(Fixed(c), Sum(Sum(a[0], Prod(a[1], a[2])), Neg(a[3]))))
    lookups = [([Advice(G + l)], [Fixed(G + 1)]) for l in range(L)]
    if spread:
        lookups.append(([Advice(G + L), Advice(G + L + 1)], [Fixed(G + 2), Fixed(G + 3)]))
    perm = [("advice", c) for c in range(A)] + [("fixed", G), ("instance", 0)]
    return plonk.ConstraintSystem(num_fixed=G + 2 + (2 if spread else 0), num_advice=A, num_instance=1, gates=gates, lookups=lookups, permutation=perm)


def halo2lib_witness(cs, k, instances, lookup_bits, groups, seed=3, num_gate_advice=15, num_lookup_advice=2):
    """A satisfying witness with full columns: `groups` chained gate groups per gate column (copy constraints d -> next b,
    first b = a constant cell, some c = public inputs, every looked-up a copied into a range-lookup column), every other
    usable row of the gate / constants columns filled with random residues (they are unconstrained), and every usable
    row of the lookup columns a random table entry. -> (fixed columns, advice columns, copies)."""
    G, L = num_gate_advice, num_lookup_advice
    spread = cs.num_advice == G + L + 2
    n = 1 << k
    usable = n - (cs.blinding_factors() + 1)
    groups = min(groups, (usable - 4) // 4)
    T = 1 << lookup_bits
    assert T < usable and len(instances) <= groups
    rng, g = random.Random(seed), np.random.default_rng(seed)

    def residues(rows):
        a = g.integers(0, 1 << 63, size=(rows, 4), dtype=np.uint64); a[:, 3] &= np.uint64((1 << 60) - 1)
        return a
    small = np.stack([plonk.fr_mont(i) for i in range(T)])                  # Montgomery forms of 0..T-1
    spread_of = lambda i: (i * i + 1) % R
    spread_mont = np.stack([plonk.fr_mont(spread_of(i)) for i in range(T)]) if spread else None
    fixed = [np.zeros((n, 4), dtype=np.uint64) for _ in range(cs.num_fixed)]
    advice = [np.zeros((n, 4), dtype=np.uint64) for _ in range(cs.num_advice)]
    one = plonk.fr_mont(1)
    fixed[G][1:usable] = residues(usable - 1)                                # constants column: free cells
    c0 = rng.randrange(R); fixed[G][0] = plonk.fr_mont(c0)
    fixed[G + 1][:T] = small                                                 # range table, zero padded
    if spread:
        fixed[G + 2][:T] = small; fixed[G + 3][:T] = spread_mont
        fixed[G + 3][T:] = 0
    A = cs.num_advice
    CONST_COL, INST_COL = A, A + 1                                           # indices in cs.permutation
    copies = []
    # lookup columns: every usable row a table entry
    look_vals = []
    for l in range(L):
        idx = g.integers(0, T, size=usable)
        advice[G + l][:usable] = small[idx]
        look_vals.append(idx)
    if spread:
        idx = g.integers(0, T, size=usable)
        advice[G + L][:usable] = small[idx]; advice[G + L + 1][:usable] = spread_mont[idx]
    next_lookup_row = [0] * L
    for c in range(G):
        col = advice[c]
        col[4 * groups:usable] = residues(usable - 4 * groups)
        prev_d = None
        for gi in range(groups):
            r = 4 * gi
            a = rng.randrange(T)
            b = c0 if gi == 0 else prev_d
            cc = instances[gi] if (c == 0 and gi < len(instances)) else rng.randrange(R)
            d = (a + b * cc) % R
            for off, v in enumerate((a, b, cc, d)):
                col[r + off] = plonk.fr_mont(v)
            fixed[c][r] = one
            copies.append(((c, r + 1), (CONST_COL, 0)) if gi == 0 else ((c, r + 1), (c, r - 1)))
            if c == 0 and gi < len(instances):
                copies.append(((c, r + 2), (INST_COL, gi)))
            if L and gi % 4 == 0:                                            # range-check a: copy it into a lookup column
                l = (c + gi) % L
                row = next_lookup_row[l]
                if row < usable:
                    advice[G + l][row] = small[a]
                    copies.append(((c, r), (G + l, row)))
                    next_lookup_row[l] = row + 1
            prev_d = d
    return fixed, advice, copies

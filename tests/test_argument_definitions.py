"""A second, independently written check of the prover-side terms the committed verifier contracts cannot exercise
(VERDICT r1 item 8): multi-set permutation arguments with their inter-set links, theta-compressed multi-column lookups and
multi-column advice -- i.e. what is specific to the sync-step shape.

tests/plonk_verifier.py checks proofs through the *verifier's* identity at one random point, and was written next to the
prover. This file does not look at proofs at all. It captures the grand-product and permuted columns the driver produces on
the CPU oracle engine (whose outputs the CUDA engine reproduces bit for bit: tests/test_gpu_plonk.py, test_gpu_lookup.py) and
checks, ROW BY ROW with Python integers, the relations that DEFINE the arguments in the halo2 book / upstream doc comments:

  permutation ([UPSTREAM] plonk/permutation/prover.rs, book "Permutation argument"), sets i = 0..m-1 over chunks of columns,
  u = usable rows, column c of the permutation carries the identity label delta^c * omega^row:
      z_0[0] = 1;   z_i[0] = z_{i-1}[u];   z_{m-1}[u] = 1
      z_i[j+1] * prod_c (v_c[j] + beta*sigma_c[j] + gamma) = z_i[j] * prod_c (v_c[j] + beta*delta^c*omega^j + gamma),  j < u
  lookup ([UPSTREAM] plonk/lookup/prover.rs, book "Lookup argument"), A / S = theta-compressed input / table:
      A = ((a_0*theta + a_1)*theta + ...) (compress_expressions folds left to right)
      A'[0:u] is a permutation of A[0:u], S'[0:u] of S[0:u];  A'[0] = S'[0];  A'[j] = S'[j] or A'[j] = A'[j-1]
      Z[0] = 1;  Z[j+1] * (A'[j]+beta)(S'[j]+gamma) = Z[j] * (A[j]+beta)(S[j]+gamma), j < u;  Z[u] = 1
A mistake shared by the prover restatement and the verifier restatement (both written from memory of upstream) cannot hide
here unless it is also a mistake in these definitions."""
import collections

import numpy as np
import pytest

from spectre_b200 import circuits, plonk
from spectre_b200.transcript import EvmTranscriptWrite
from tests.plonk_oracle_engine import OracleEngine, SeededRng

R = plonk.R_MOD


class RecordingEngine(OracleEngine):
    def __init__(self, k, j):
        super().__init__(k, j)
        self.perm, self.pairs, self.prods, self.theta = [], [], [], None

    def graph_evaluate(self, p, fixed, advice, instance, beta, gamma, theta, y, values, size, rot_scale):
        if self.theta is None and np.any(theta):
            self.theta = np.array(theta, dtype=np.uint64).copy()     # the first non-zero theta is the lookup compression's
        super().graph_evaluate(p, fixed, advice, instance, beta, gamma, theta, y, values, size, rot_scale)

    def permutation_product(self, values, sigma, first_col, beta, gamma, blinds, last_z, z):
        out = super().permutation_product(values, sigma, first_col, beta, gamma, blinds, last_z, z)
        self.perm.append(dict(values=[b.a.copy() for b in values], sigma=[b.a.copy() for b in sigma], first_col=first_col, beta=beta.copy(), gamma=gamma.copy(), z=z.a.copy()))
        return out

    def permute_expression_pair(self, a, s, usable, out_a, out_s):
        super().permute_expression_pair(a, s, usable, out_a, out_s)
        self.pairs.append(dict(a=a.a.copy(), s=s.a.copy(), pa=out_a.a.copy(), ps=out_s.a.copy(), usable=usable))

    def lookup_product(self, ci, ct, pi, pt, beta, gamma, blinds, z):
        super().lookup_product(ci, ct, pi, pt, beta, gamma, blinds, z)
        self.prods.append(dict(a=ci.a.copy(), s=ct.a.copy(), pa=pi.a.copy(), ps=pt.a.copy(), beta=beta.copy(), gamma=gamma.copy(), z=z.a.copy()))


def ints(orc, arr):
    return orc.fr_ints(np.ascontiguousarray(arr).reshape(-1, 4))


@pytest.mark.parametrize("shape", ["halo2lib_spread", "wide"])
def test_arguments_satisfy_their_definitions_row_by_row(orc, shape):
    k, instances = 8, [5, 6, 7]
    n = 1 << k
    if shape == "halo2lib_spread":      # 3 gate columns, 2 range-lookup columns, the two-column spread lookup: 9 permutation columns
        cs = circuits.halo2lib_shape(3, 2)
        fixed, adv, copies = circuits.halo2lib_witness(cs, k, instances, lookup_bits=4, groups=20, num_gate_advice=3, num_lookup_advice=2)
    else:
        cs = circuits.wide_shape(3)
        fixed, adv, copies = circuits.wide_witness(cs, k, instances, lookup_bits=4, groups=20)
    E = RecordingEngine(k, cs.degree())
    pk = plonk.keygen(E, cs, k, fixed, copies)
    T = EvmTranscriptWrite(pk.vk_digest)
    plonk.create_proof(E, pk, [instances], adv, SeededRng(11), T)
    u = pk.usable_rows
    w = plonk.omega_of(k)
    chunk = cs.chunk_len()
    n_cols = len(cs.permutation)
    assert len(E.perm) == -(-n_cols // chunk) and (shape != "halo2lib_spread" or len(E.perm) >= 3)      # several sets: the inter-set links are exercised

    # ---- permutation argument ----
    prev_last = 1
    for i, rec in enumerate(E.perm):
        beta, gamma = ints(orc, rec["beta"])[0], ints(orc, rec["gamma"])[0]
        z = ints(orc, rec["z"])
        vals = [ints(orc, v) for v in rec["values"]]
        sig = [ints(orc, s) for s in rec["sigma"]]
        assert rec["first_col"] == i * chunk
        assert z[0] == prev_last, "set %d does not start where set %d ended" % (i, i - 1)
        for j in range(u):
            left, right = z[j + 1], z[j]
            for c in range(len(vals)):
                left = left * (vals[c][j] + beta * sig[c][j] + gamma) % R
                right = right * (vals[c][j] + beta * pow(plonk.DELTA, rec["first_col"] + c, R) * pow(w, j, R) + gamma) % R
            assert left == right, "permutation set %d breaks its recurrence at row %d" % (i, j)
        prev_last = z[u]
    assert prev_last == 1, "the last set's product over the usable rows is not 1"
    # sigma really is a permutation of the identity labels delta^c * omega^row
    labels = collections.Counter()
    for rec in E.perm:
        for sg in rec["sigma"]:
            labels.update(ints(orc, sg))
    ident = collections.Counter(pow(plonk.DELTA, c, R) * pow(w, j, R) % R for c in range(n_cols) for j in range(n))
    assert labels == ident

    # ---- lookup arguments ----
    assert len(E.pairs) == len(cs.lookups) == len(E.prods)
    for li, (pair, prod) in enumerate(zip(E.pairs, E.prods)):
        A, S = ints(orc, pair["a"]), ints(orc, pair["s"])
        PA, PS = ints(orc, pair["pa"]), ints(orc, pair["ps"])
        assert pair["usable"] == u
        assert collections.Counter(PA[:u]) == collections.Counter(A[:u]), "A' is not a permutation of A (lookup %d)" % li
        assert collections.Counter(PS[:u]) == collections.Counter(S[:u]), "S' is not a permutation of S (lookup %d)" % li
        assert PA[0] == PS[0]
        for j in range(1, u):
            assert PA[j] == PS[j] or PA[j] == PA[j - 1], "lookup %d row %d: A' is neither S' nor the previous A'" % (li, j)
        beta, gamma = ints(orc, prod["beta"])[0], ints(orc, prod["gamma"])[0]
        Z = ints(orc, prod["z"])
        assert ints(orc, prod["a"]) == A and ints(orc, prod["pa"])[:u] == PA[:u]
        PAb, PSb = ints(orc, prod["pa"]), ints(orc, prod["ps"])     # blinded rows included (not used below u)
        assert Z[0] == 1
        for j in range(u):
            assert Z[j + 1] * (PAb[j] + beta) % R * (PSb[j] + gamma) % R == Z[j] * (A[j] + beta) % R * (S[j] + gamma) % R, "lookup %d product breaks at row %d" % (li, j)
        assert Z[u] == 1
    # theta-compression: A = compress_expressions(inputs) = ((e_0 * theta + e_1) * theta + ...), every expression evaluated here
    # from the raw columns with Python integers; same for the table side
    theta = ints(orc, E.theta)[0]
    advs = adv if isinstance(adv, list) else [adv]
    col = {"fixed": [ints(orc, c) for c in fixed], "advice": [ints(orc, c) for c in advs]}
    inst_col = [0] * n
    inst_col[:len(instances)] = instances
    col["instance"] = [inst_col]

    def ev(e, j):
        t = e[0]
        if t == "const": return e[1]
        if t in ("fixed", "advice", "instance"): return col[t][e[1]][(j + e[2]) % n]
        if t == "neg": return -ev(e[1], j) % R
        if t == "scaled": return ev(e[1], j) * e[2] % R
        if t == "sum": return (ev(e[1], j) + ev(e[2], j)) % R
        return ev(e[1], j) * ev(e[2], j) % R
    assert any(len(ins) > 1 for ins, _ in cs.lookups), "a multi-column (theta-compressed) lookup is part of both shapes"
    for li, (ins, tbs) in enumerate(cs.lookups):
        A, S = ints(orc, E.pairs[li]["a"]), ints(orc, E.pairs[li]["s"])
        for j in range(u):
            acc_a = acc_s = 0
            for e in ins: acc_a = (acc_a * theta + ev(e, j)) % R
            for e in tbs: acc_s = (acc_s * theta + ev(e, j)) % R
            assert A[j] == acc_a and S[j] == acc_s, "lookup %d: theta-compression differs from compress_expressions at row %d" % (li, j)

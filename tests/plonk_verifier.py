"""An independent verifier for the proofs spectre_b200/plonk.py produces -- TEST INFRASTRUCTURE, pure Python integers.

It restates halo2's verifier ([UPSTREAM] halo2_proofs/src/plonk/verifier.rs, plonk/{permutation,lookup,vanishing}/
verifier.rs, poly/kzg/multiopen/shplonk/verifier.rs) with one substitution: the final pairing check
e(A, [tau]_2) = e(B, [1]_2) is decided in G1 as tau * A == B, which is sound for the test SRS whose tau is known
(halo2-base's gen_srs seeds it with zeros; reference call sites prover/src/cli.rs:48). The reference's own tests check
validity the same way (`test_step_proofgen`, lightclient-circuits/src/sync_step_circuit.rs:481-503: prove, then verify).
The same identity, term by term, is what the committed verifier contract evaluates (sync_step_verifier.sol:507-590);
tests/yul_harness.py replays that contract itself on a K = 23 proof.
"""
from spectre_b200 import plonk
from spectre_b200.transcript import EvmTranscriptRead
from tests import pyref

R = pyref.R_MOD
G1 = (1, 2)


def _inv(a):
    return pow(a % R, -1, R)


def evaluate(e, fixed, advice, instance):
    """expression tree at a point, given evaluations keyed by (col, rot)"""
    t = e[0]
    if t == "const": return e[1]
    if t == "fixed": return fixed[(e[1], e[2])]
    if t == "advice": return advice[(e[1], e[2])]
    if t == "instance": return instance[(e[1], e[2])]
    if t == "neg": return -evaluate(e[1], fixed, advice, instance) % R
    if t == "sum": return (evaluate(e[1], fixed, advice, instance) + evaluate(e[2], fixed, advice, instance)) % R
    if t == "prod": return evaluate(e[1], fixed, advice, instance) * evaluate(e[2], fixed, advice, instance) % R
    if t == "scaled": return evaluate(e[1], fixed, advice, instance) * e[2] % R
    raise ValueError(t)


def lagrange_evals(k, x, rows):
    """l_i(x) for the given rows of the size-2^k domain"""
    n = 1 << k
    w = pyref.omega(k)
    xn = pow(x, n, R)
    out = {}
    for i in rows:
        wi = pow(w, i % n, R)
        out[i] = (xn - 1) * wi % R * _inv(n * (x - wi)) % R
    return out


def verify(cs, k, vk_digest, fixed_commitments, sigma_commitments, instances, proof, tau, transcript_read=EvmTranscriptRead, trace=None):
    """True / raises AssertionError with the failing check. transcript_read: the reading transcript class (Keccak EVM transcript by
    default, spectre_b200.poseidon.PoseidonTranscriptRead for proofs made over the Poseidon transcript). trace: a dict that receives
    the intermediate values (challenges, x^n, Lagrange terms, the expected quotient evaluation, the opening-set quantities) so that
    they can be compared with what the reference's verifier contract computes for the same proof."""
    n = 1 << k
    bf = cs.blinding_factors()
    usable = n - (bf + 1)
    w = pyref.omega(k)
    T = transcript_read(vk_digest, proof)
    for col in instances:
        for v in col:
            T.common_scalar(v)
    advice_c = [T.read_ec_point() for _ in range(cs.num_advice)]
    theta = T.squeeze_challenge()
    permuted_c = [(T.read_ec_point(), T.read_ec_point()) for _ in cs.lookups]
    beta = T.squeeze_challenge(); gamma = T.squeeze_challenge()
    chunk = max(1, cs.chunk_len())
    n_sets = -(-len(cs.permutation) // chunk) if cs.permutation else 0
    perm_c = [T.read_ec_point() for _ in range(n_sets)]
    lookz_c = [T.read_ec_point() for _ in cs.lookups]
    random_c = T.read_ec_point()
    y = T.squeeze_challenge()
    h_c = [T.read_ec_point() for _ in range(cs.degree() - 1)]
    x = T.squeeze_challenge()
    adv = {q: T.read_scalar() for q in cs.advice_queries}
    fix = {q: T.read_scalar() for q in cs.fixed_queries}
    random_eval = T.read_scalar()
    sigma_evals = [T.read_scalar() for _ in cs.permutation]
    perm_evals = []
    for s in range(n_sets):
        e0, e1 = T.read_scalar(), T.read_scalar()
        perm_evals.append((e0, e1, T.read_scalar() if s + 1 < n_sets else None))
    look_evals = [tuple(T.read_scalar() for _ in range(5)) for _ in cs.lookups]   # z, z_next, a', a'_inv, s'

    # ---- the quotient identity at x ----
    xn = pow(x, n, R)
    blind_rows = range(usable + 1, n)
    max_inst = max([len(c) for c in instances] or [0])
    inst_rots = sorted({r for _, r in cs.instance_queries})
    L = lagrange_evals(k, x, [0, usable] + list(blind_rows) + [i - r for r in inst_rots for i in range(max_inst)])
    l0, l_last = L[0], L[usable]
    l_blind = sum(L[i] for i in blind_rows) % R
    l_active = (1 - l_last - l_blind) % R
    # instance evaluations from the public inputs: inst(x * w^r) = sum_i inst[i] * l_i(x w^r) = sum_i inst[i] * l_{i-r}(x)
    inst = {(c, r): sum(v * L[i - r] for i, v in enumerate(instances[c])) % R for c, r in cs.instance_queries}
    acc = 0
    for g in cs.gates:
        acc = (acc * y + evaluate(g, fix, adv, inst)) % R
    if n_sets:
        col_eval = lambda kind, c: {"fixed": fix, "advice": adv, "instance": inst}[kind][(c, 0)]
        acc = (acc * y + l0 * (1 - perm_evals[0][0])) % R
        zl = perm_evals[-1][0]
        acc = (acc * y + l_last * (zl * zl - zl)) % R
        for s in range(1, n_sets):
            acc = (acc * y + l0 * (perm_evals[s][0] - perm_evals[s - 1][2])) % R
        for s in range(n_sets):
            lo, hi = s * chunk, min((s + 1) * chunk, len(cs.permutation))
            left, right = perm_evals[s][1], perm_evals[s][0]
            for c in range(lo, hi):
                kind, col = cs.permutation[c]
                left = left * (col_eval(kind, col) + beta * sigma_evals[c] + gamma) % R
                right = right * (col_eval(kind, col) + beta * x % R * pow(plonk.DELTA, c, R) + gamma) % R
            acc = (acc * y + l_active * (left - right)) % R
    for (ins, tbs), (z, z_next, a_p, a_inv, s_p) in zip(cs.lookups, look_evals):
        ci = 0
        for e in ins: ci = (ci * theta + evaluate(e, fix, adv, inst)) % R
        ct = 0
        for e in tbs: ct = (ct * theta + evaluate(e, fix, adv, inst)) % R
        acc = (acc * y + l0 * (1 - z)) % R
        acc = (acc * y + l_last * (z * z - z)) % R
        acc = (acc * y + l_active * (z_next * (a_p + beta) % R * (s_p + gamma) - z * (ci + beta) % R * (ct + gamma))) % R
        acc = (acc * y + l0 * (a_p - s_p)) % R
        acc = (acc * y + l_active * (a_p - s_p) % R * (a_p - a_inv)) % R
    expected_h = acc * _inv(xn - 1) % R
    if trace is not None:
        trace.update(theta=theta, beta=beta, gamma=gamma, y=y, x=x, x_n=xn, l_0=l0, l_last=l_last, l_blind=l_blind,
                     quotient_numerator=acc, expected_h=expected_h, instance_evals=dict(inst))

    # ---- multi-open: queries in the prover's order, with the verifier's commitments ----
    ec_add, ec_mul = pyref.ec_add, pyref.ec_mul
    h_commit = None
    for c in reversed(h_c):                                   # sum_i x^(n i) H_i
        h_commit = ec_add(ec_mul(h_commit, xn) if h_commit else None, c)
    xw = lambda r: x * pow(w, r % n, R) % R
    q = []
    for (c, r) in cs.advice_queries: q.append((("advice", c), advice_c[c], xw(r), adv[(c, r)]))
    for s, (e0, e1, _) in enumerate(perm_evals):
        q.append((("perm", s), perm_c[s], x, e0)); q.append((("perm", s), perm_c[s], xw(1), e1))
    for s in reversed(range(n_sets - 1)):
        q.append((("perm", s), perm_c[s], xw(-(bf + 1)), perm_evals[s][2]))
    for li, (z, z_next, a_p, a_inv, s_p) in enumerate(look_evals):
        pin, ptab = permuted_c[li]
        q += [(("lk_z", li), lookz_c[li], x, z), (("lk_a", li), pin, x, a_p), (("lk_s", li), ptab, x, s_p), (("lk_a", li), pin, xw(-1), a_inv), (("lk_z", li), lookz_c[li], xw(1), z_next)]
    for (c, r) in cs.fixed_queries: q.append((("fixed", c), fixed_commitments[c], xw(r), fix[(c, r)]))
    for c, e in enumerate(sigma_evals): q.append((("sigma", c), sigma_commitments[c], x, e))
    q.append((("h",), h_commit, x, expected_h)); q.append((("random",), random_c, x, random_eval))
    commits = {pid: c for pid, c, _, _ in q}
    sets = plonk.rotation_sets([(pid, pt, ev) for pid, _, pt, ev in q])
    y2 = T.squeeze_challenge(); v = T.squeeze_challenge()
    h1 = T.read_ec_point()
    u = T.squeeze_challenge()
    if trace is not None:
        trace.update(shplonk_y=y2, shplonk_v=v, shplonk_u=u)
    h2 = T.read_ec_point()
    assert T.pos == len(proof), "trailing bytes in proof"
    super_pts = []
    for pts, _, _ in sets:
        for p in pts:
            if p not in super_pts: super_pts.append(p)

    def interp_eval(pts, evs, at):                            # Lagrange interpolation of (pts, evs), evaluated at `at`
        tot = 0
        for j, (pj, ej) in enumerate(zip(pts, evs)):
            num = den = 1
            for m, pm in enumerate(pts):
                if m != j:
                    num = num * (at - pm) % R; den = den * (pj - pm) % R
            tot = (tot + ej * num % R * _inv(den)) % R
        return tot
    acc_pt, acc_const, z0 = None, 0, None
    for i, (pts, pids, evs) in enumerate(sets):
        zi = 1
        for p in super_pts:
            if p not in pts: zi = zi * (u - p) % R
        if i == 0: z0 = zi
        outer = pow(v, i, R) * zi % R
        for j, pid in enumerate(pids):
            wij = outer * pow(y2, j, R) % R
            acc_pt = ec_add(acc_pt, ec_mul(commits[pid], wij))
            acc_const = (acc_const + wij * interp_eval(pts, evs[j], u)) % R
    zt = 1
    for p in super_pts: zt = zt * (u - p) % R
    rhs = ec_add(acc_pt, ec_mul(G1, (-acc_const) % R))
    rhs = ec_add(rhs, ec_mul(h1, (-zt) % R))
    rhs = ec_mul(rhs, _inv(z0)) if rhs else None
    lhs = ec_mul(h2, (tau - u) % R)
    assert lhs == rhs, "SHPLONK opening check failed (tau * A != B)"
    return True

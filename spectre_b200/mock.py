"""MockProver -- host-side constraint-satisfaction check, the mirror of halo2_proofs::dev::MockProver that the
reference's unit tests are built on (`MockProver::run(k, &circuit, instances).assert_satisfied()`,
lightclient-circuits/src/sync_step_circuit.rs:459-479, committee_update_circuit.rs:313-333). No MSM, no NTT, no GPU:
plain integers over the usable rows. Meant for small k (tests, witness debugging), like upstream's."""
from . import plonk

R = plonk.R_MOD


def _ints(col):
    return [plonk.fr_int(row) for row in col]


def _eval(e, row, n, fixed, advice, instance):
    t = e[0]
    if t == "const": return e[1]
    if t == "fixed": return fixed[e[1]][(row + e[2]) % n]
    if t == "advice": return advice[e[1]][(row + e[2]) % n]
    if t == "instance": return instance[e[1]][(row + e[2]) % n]
    if t == "neg": return -_eval(e[1], row, n, fixed, advice, instance) % R
    if t == "sum": return (_eval(e[1], row, n, fixed, advice, instance) + _eval(e[2], row, n, fixed, advice, instance)) % R
    if t == "prod": return _eval(e[1], row, n, fixed, advice, instance) * _eval(e[2], row, n, fixed, advice, instance) % R
    if t == "scaled": return _eval(e[1], row, n, fixed, advice, instance) * e[2] % R
    raise ValueError(t)


class VerifyFailure(Exception):
    pass


def run(cs, k, fixed_columns, advice_columns, instances, copies):
    """-> list of failure strings (empty = satisfied). Columns: (n, 4) Montgomery arrays as create_proof takes them;
    instances: per instance column a list of ints; copies: ((perm column, row), (perm column, row)) equalities."""
    n = 1 << k
    usable = n - (cs.blinding_factors() + 1)
    fixed = [_ints(c) for c in fixed_columns]
    advice = [_ints(c) for c in advice_columns]
    inst = [list(col) + [0] * (n - len(col)) for col in instances]
    failures = []
    for gi, g in enumerate(cs.gates):
        for row in range(usable):
            if _eval(g, row, n, fixed, advice, inst) != 0:
                failures.append("gate %d not satisfied on row %d" % (gi, row))
    for li, (ins, tbs) in enumerate(cs.lookups):
        table = {tuple(_eval(e, row, n, fixed, advice, inst) for e in tbs) for row in range(usable)}
        for row in range(usable):
            if tuple(_eval(e, row, n, fixed, advice, inst) for e in ins) not in table:
                failures.append("lookup %d: input on row %d is not in the table" % (li, row))
    cols = {"fixed": fixed, "advice": advice, "instance": inst}
    cell = lambda c, r: cols[cs.permutation[c][0]][cs.permutation[c][1]][r]
    for (c1, r1), (c2, r2) in copies:
        if cell(c1, r1) != cell(c2, r2):
            failures.append("copy constraint (%d, %d) = (%d, %d) not satisfied" % (c1, r1, c2, r2))
    return failures


def assert_satisfied(cs, k, fixed_columns, advice_columns, instances, copies):
    failures = run(cs, k, fixed_columns, advice_columns, instances, copies)
    if failures:
        raise VerifyFailure("; ".join(failures[:8]) + (" ... (%d failures)" % len(failures) if len(failures) > 8 else ""))

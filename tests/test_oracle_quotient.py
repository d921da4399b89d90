"""CPU: the oracle's evaluate_h restatement against plain-Python integer arithmetic on tiny systems."""
import numpy as np
import pytest

from tests import pyref
from tests.quotient_common import ADD, MUL, HORNER, K_ADVICE, K_FIXED, K_INTER, K_PREV, K_Y, K_CONST

R = pyref.R_MOD


def test_graph_small_program(orc):
    size, rot_scale = 16, 2
    a = orc.fr_random_chacha(size, 1); f = orc.fr_random_chacha(size, 2); prev = orc.fr_random_chacha(size, 3)
    consts = orc.fr([5]); y = orc.fr_random_chacha(1, 4)[0]
    rotations = np.array([0, -1], dtype=np.int32)
    prog = np.array([MUL, 0, K_ADVICE, 0, K_FIXED, 0 | (1 << 16),      # t0 = a[i] * f[i - rot_scale]
                     ADD, 1, K_INTER, 0, K_CONST, 0,                   # t1 = t0 + 5
                     HORNER | (1 << 8), 2, K_PREV, 0, K_Y, 0, K_INTER, 1], dtype=np.uint32)
    zero = np.zeros((1, 4), dtype=np.uint64)
    bgty = np.stack([zero[0], zero[0], zero[0], y])
    got = orc.fr_ints(orc.graph_evaluate(prog, 3, 3, consts, rotations, [f], [a], [], zero, bgty, prev, rot_scale))
    ai, fi, pi, yi = orc.fr_ints(a), orc.fr_ints(f), orc.fr_ints(prev), orc.fr_ints(y)[0]
    assert got == [(pi[i] * yi + ai[i] * fi[(i - rot_scale) % size] + 5) % R for i in range(size)]


def test_permutation_and_lookup_terms(orc):
    size, rot_scale = 16, 2
    mk = lambda s: orc.fr_random_chacha(size, s)
    z, cv, sg = [mk(1), mk(2)], [mk(3), mk(4), mk(5)], [mk(6), mk(7), mk(8)]
    l0, ll, la, v0 = mk(9), mk(10), mk(11), mk(12)
    beta, gamma, y = orc.fr_random_chacha(3, 13)
    wext = orc.fr([pyref.omega(4)])[0]
    last_rot = -3
    got = orc.fr_ints(orc.permutation_constraints(v0, rot_scale, last_rot, 2, z, cv, sg, l0, ll, la, beta, gamma, y, wext))
    I = orc.fr_ints
    zi, ci, si = [I(t) for t in z], [I(t) for t in cv], [I(t) for t in sg]
    l0i, lli, lai, vi = I(l0), I(ll), I(la), I(v0)
    b, g, yy = I(beta)[0], I(gamma)[0], I(y)[0]
    delta = pow(7, 1 << 28, R); w = pyref.omega(4)
    want = []
    for i in range(size):
        rn, rl = (i + rot_scale) % size, (i + last_rot * rot_scale) % size
        v = vi[i]
        v = (v * yy + (1 - zi[0][i]) * l0i[i]) % R
        v = (v * yy + (zi[1][i] ** 2 - zi[1][i]) * lli[i]) % R
        v = (v * yy + (zi[1][i] - zi[0][rl]) * l0i[i]) % R
        cur = b * pyref.ZETA * pow(w, i, R) % R
        for s, (lo, hi) in enumerate(((0, 2), (2, 3))):
            left, right = zi[s][rn], zi[s][i]
            for c in range(lo, hi):
                left = left * (ci[c][i] + b * si[c][i] + g) % R
            for c in range(lo, hi):
                right = right * (ci[c][i] + cur + g) % R
                cur = cur * delta % R
            v = (v * yy + (left - right) * lai[i]) % R
        want.append(v)
    assert got == want
    prod, pin, ptb, tv = mk(20), mk(21), mk(22), mk(23)
    got = I(orc.lookup_constraints(v0, rot_scale, prod, pin, ptb, tv, l0, ll, la, beta, gamma, y))
    p, a, t, tvi = I(prod), I(pin), I(ptb), I(tv)
    want = []
    for i in range(size):
        rn, rp = (i + rot_scale) % size, (i - rot_scale) % size
        v = vi[i]
        v = (v * yy + (1 - p[i]) * l0i[i]) % R
        v = (v * yy + (p[i] ** 2 - p[i]) * lli[i]) % R
        v = (v * yy + (p[rn] * (a[i] + b) * (t[i] + g) - p[i] * tvi[i]) * lai[i]) % R
        v = (v * yy + (a[i] - t[i]) * l0i[i]) % R
        v = (v * yy + (a[i] - t[i]) * (a[i] - a[rp]) * lai[i]) % R
        want.append(v)
    assert got == want


def _schedule(prog, ncalc):
    """libspectre_b200's host-side scheduling pass (no device needed) -> (words, slots, calculations)"""
    import ctypes
    from spectre_b200 import halo2
    lib = halo2.load_library()
    prog = np.ascontiguousarray(prog, dtype=np.uint32)
    out = np.zeros(4 * len(prog) + 64, dtype=np.uint32)
    cnt, slots, nc = ctypes.c_size_t(0), ctypes.c_uint32(0), ctypes.c_uint32(0)
    rc = lib.spb_test_schedule_program(prog.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(prog)), ctypes.c_uint32(ncalc), out.ctypes.data_as(ctypes.c_void_p),
                                       ctypes.c_size_t(len(out)), ctypes.byref(cnt), ctypes.byref(slots), ctypes.byref(nc))
    assert rc == 0, rc
    return out[:cnt.value].copy(), slots.value, nc.value


def test_scheduled_program_is_the_same_function_with_few_live_intermediates(orc):
    """spb_graph_evaluate_dev reschedules the flat GraphEvaluator program before running it (Horner split into one-part steps next
    to their producers, intermediates renamed to slots by liveness: csrc/quotient.cu). The rescheduled words, run through the
    oracle's interpreter, give the same value on every row as the original program; the 15-gate halo2-lib shape drops from 76
    intermediates to a handful of slots, the lookup-compression and lookup-value programs keep working."""
    from spectre_b200 import circuits
    cs = circuits.halo2lib_shape()
    n = 64
    fixed = [orc.fr_random_chacha(n, 300 + i) for i in range(cs.num_fixed)]
    advice = [orc.fr_random_chacha(n, 400 + i) for i in range(cs.num_advice)]
    inst = [orc.fr_random_chacha(n, 500)]
    bgty = orc.fr_random_chacha(4, 600)
    prev = orc.fr_random_chacha(n, 700)
    ch = np.zeros((1, 4), np.uint64)
    progs = [cs.gates_program(), cs.lookup_value_program(len(cs.lookups) - 1), cs.lookup_compress_program(cs.lookups[-1][0])]
    for p in progs:
        words, slots, nc = _schedule(p["prog"], p["ncalc"])
        assert slots <= 8 and slots <= p["ncalc"]
        want = orc.graph_evaluate(p["prog"], p["ncalc"], p["ncalc"], p["constants"], p["rotations"], fixed, advice, inst, ch, bgty, prev, 1)
        got = orc.graph_evaluate(words, nc, slots, p["constants"], p["rotations"], fixed, advice, inst, ch, bgty, prev, 1)
        assert np.array_equal(got, want)
    assert progs[0]["ncalc"] >= 60 and _schedule(progs[0]["prog"], progs[0]["ncalc"])[1] <= 6
    # a program that writes one target twice is left alone (the pass declines)
    import ctypes
    from spectre_b200 import halo2
    twice = np.array([0, 0, 3, 0, 3, 1 << 16, 0, 0, 1, 0, 3, 0], dtype=np.uint32)      # t0 = a0 + a0'; t0 = t0 + a0
    out = np.zeros(64, np.uint32); cnt, slots, nc = ctypes.c_size_t(0), ctypes.c_uint32(0), ctypes.c_uint32(0)
    assert halo2.load_library().spb_test_schedule_program(twice.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(twice)), ctypes.c_uint32(2),
                                                          out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(64), ctypes.byref(cnt), ctypes.byref(slots), ctypes.byref(nc)) != 0


@pytest.mark.parametrize("seed", range(40))
def test_scheduler_on_random_programs(orc, seed):
    """Random flat programs (every op code, Horners of 0..5 parts whose parts are produced out of order, repeated sources,
    PreviousValue, dead calculations, results that are never read): the rescheduled words evaluate to the same value on every row
    as the original ones, use no more slots than there were intermediates, and the row's result is still the last calculation's."""
    rng = np.random.default_rng(1000 + seed)
    n, nf, na = 32, 3, 4
    fixed = [orc.fr_random_chacha(n, 800 + i) for i in range(nf)]
    advice = [orc.fr_random_chacha(n, 810 + i) for i in range(na)]
    inst = [orc.fr_random_chacha(n, 820)]
    constants = orc.fr([0, 1, 7, 2 ** 200 + 5])
    rotations = np.array([0, 1, -1, 3], dtype=np.int32)
    bgty = orc.fr_random_chacha(4, 830 + seed)
    prev = orc.fr_random_chacha(n, 840 + seed)
    ch = orc.fr_random_chacha(2, 850)
    ncalc = int(rng.integers(1, 40))
    words = []

    def source(c):
        kind = int(rng.choice([0, 1, 1, 1, 2, 3, 3, 4, 5, 6, 7, 8, 9, 10])) if c else int(rng.choice([0, 2, 3, 4, 5, 6, 9, 10]))
        if kind == 0: return [0, int(rng.integers(0, 4))]
        if kind == 1: return [1, int(rng.integers(0, c))]
        if kind == 2: return [2, int(rng.integers(0, nf)) | int(rng.integers(0, 4)) << 16]
        if kind == 3: return [3, int(rng.integers(0, na)) | int(rng.integers(0, 4)) << 16]
        if kind == 4: return [4, 0 | int(rng.integers(0, 4)) << 16]
        if kind == 5: return [5, int(rng.integers(0, 2))]
        return [kind, 0]
    for c in range(ncalc):
        op = int(rng.choice([0, 1, 2, 2, 3, 4, 5, 6, 6, 7]))
        if op <= 2:
            words += [op, c] + source(c) + source(c)
        elif op == 6:
            parts = int(rng.integers(0, 6))
            words += [6 | parts << 8, c] + source(c) + source(c)
            for _ in range(parts):
                words += source(c)
        else:
            words += [op, c] + source(c)
    prog = np.array(words, dtype=np.uint32)
    new_words, slots, nc = _schedule(prog, ncalc)
    assert slots <= ncalc and nc >= ncalc
    want = orc.graph_evaluate(prog, ncalc, ncalc, constants, rotations, fixed, advice, inst, ch, bgty, prev, 1)
    got = orc.graph_evaluate(new_words, nc, slots, constants, rotations, fixed, advice, inst, ch, bgty, prev, 1)
    assert np.array_equal(got, want)

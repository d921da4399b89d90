"""CPU: the oracle's evaluate_h restatement against plain-Python integer arithmetic on tiny systems."""
import numpy as np

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

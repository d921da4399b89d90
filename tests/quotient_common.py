"""Synthetic constraint systems for the quotient-numerator tests (graph programs in the flat encoding of
include/spectre_b200.h / oracle/halo2_oracle.c)."""
import random

import numpy as np

ADD, SUB, MUL, SQUARE, DOUBLE, NEGATE, HORNER, STORE = range(8)
K_CONST, K_INTER, K_FIXED, K_ADVICE, K_INSTANCE, K_CHALLENGE, K_BETA, K_GAMMA, K_THETA, K_Y, K_PREV = range(11)


def random_program(rng, ncalc, n_const, n_rot, n_fixed, n_advice, n_instance, n_chal):
    """Every calculation writes a fresh intermediate and may read any earlier one (as GraphEvaluator does)."""
    words = []

    def src(n_done):
        kinds = [K_CONST, K_FIXED, K_ADVICE, K_INSTANCE, K_CHALLENGE, K_BETA, K_GAMMA, K_THETA, K_Y, K_PREV] + ([K_INTER] * 4 if n_done else [])
        kind = rng.choice(kinds)
        idx, rot = 0, 0
        if kind == K_CONST: idx = rng.randrange(n_const)
        elif kind == K_INTER: idx = rng.randrange(n_done)
        elif kind == K_FIXED: idx, rot = rng.randrange(n_fixed), rng.randrange(n_rot)
        elif kind == K_ADVICE: idx, rot = rng.randrange(n_advice), rng.randrange(n_rot)
        elif kind == K_INSTANCE: idx, rot = rng.randrange(n_instance), rng.randrange(n_rot)
        elif kind == K_CHALLENGE: idx = rng.randrange(n_chal)
        return [kind, idx | (rot << 16)]

    for c in range(ncalc):
        op = rng.choice([ADD, SUB, MUL, MUL, SQUARE, DOUBLE, NEGATE, HORNER, STORE])
        if op in (ADD, SUB, MUL):
            words += [op, c] + src(c) + src(c)
        elif op == HORNER:
            nparts = rng.randrange(0, 5)
            words += [op | (nparts << 8), c] + src(c) + src(c)
            for _ in range(nparts):
                words += src(c)
        else:
            words += [op, c] + src(c)
    return np.array(words, dtype=np.uint32)

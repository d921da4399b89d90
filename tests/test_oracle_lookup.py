"""CPU: the oracle's permute_expression_pair against a direct Python transcription of the upstream algorithm."""
import random

import numpy as np
import pytest

from tests import pyref


def py_permute(inp, tab):
    usable = len(inp)
    pin = sorted(inp)
    leftover = {}
    for v in tab:
        leftover[v] = leftover.get(v, 0) + 1
    ptab = [0] * usable
    repeated = []
    for row, v in enumerate(pin):
        if row == 0 or v != pin[row - 1]:
            ptab[row] = v
            assert leftover.get(v, 0) > 0
            leftover[v] -= 1
        else:
            repeated.append(row)
    for v in sorted(leftover):
        for _ in range(leftover[v]):
            ptab[repeated.pop()] = v
    assert not repeated
    return pin, ptab


@pytest.mark.parametrize("seed,n,table_size", [(1, 1, 1), (2, 37, 5), (3, 200, 16), (4, 500, 500), (5, 1000, 64)])
def test_permute_matches_python(orc, seed, n, table_size):
    rng = random.Random(seed)
    vals = [rng.randrange(pyref.R_MOD) for _ in range(table_size)] + [0]
    # table: every value at least once, padded with repeats to n rows; input: values drawn from the table
    tab = (vals * (n // len(vals) + 1))[:n] if n >= len(vals) else vals[:n]
    present = sorted(set(tab))
    inp = [rng.choice(present) for _ in range(n)]
    pi, pt = orc.permute_expression_pair(orc.fr(inp), orc.fr(tab))
    wi, wt = py_permute(inp, tab)
    assert orc.fr_ints(pi) == wi and orc.fr_ints(pt) == wt
    # the two properties the lookup argument needs: multisets preserved; each row a' == s' or a' == a'[row-1]
    assert sorted(orc.fr_ints(pt)) == sorted(tab)
    a, s = orc.fr_ints(pi), orc.fr_ints(pt)
    assert all(a[i] == s[i] or (i > 0 and a[i] == a[i - 1]) for i in range(n))


def test_missing_value_fails(orc):
    with pytest.raises(ValueError):
        orc.permute_expression_pair(orc.fr([1, 2, 3]), orc.fr([1, 2, 4]))

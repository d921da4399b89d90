"""GPU parity of the field layer (spb_test_field_op through the C ABI) against the oracle / Python ints."""
import random

import numpy as np
import pytest

from tests import pyref
from tests.gpu_common import be  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("field", ["fr", "fq"])
def test_field_ops_bit_exact(be, orc, field):
    mod = pyref.R_MOD if field == "fr" else pyref.P_MOD
    rng = random.Random(42)
    edge = [0, 1, 2, mod - 1, mod - 2, (mod - 1) // 2, 1 << 253, (1 << 64) - 1, 1 << 64, (1 << 128) - 1, 1 << 224]
    a = edge + [rng.randrange(mod) for _ in range(5000)]
    b = [rng.randrange(mod) for _ in range(5000)] + list(reversed(edge))
    A = orc.to_mont(a, mod); B = orc.to_mont(b, mod)
    assert orc.from_mont(be.test_field_op(field, "mul", A, B), mod) == [x * y % mod for x, y in zip(a, b)]
    assert orc.from_mont(be.test_field_op(field, "add", A, B), mod) == [(x + y) % mod for x, y in zip(a, b)]
    assert orc.from_mont(be.test_field_op(field, "sub", A, B), mod) == [(x - y) % mod for x, y in zip(a, b)]


def test_mul_raw_limb_patterns(be):
    """worst-case carry patterns, operands given as raw limbs (not converted to Montgomery form)."""
    for mod, f in ((pyref.R_MOD, "fr"), (pyref.P_MOD, "fq")):
        vals = [mod - 1 - i for i in range(16)] + [((1 << 254) - 1) % mod, (1 << 224) - 1, mod >> 1, (mod >> 1) + 1]
        a = [x for x in vals for _ in vals]; b = [y for _ in vals for y in vals]
        A = np.array([[(x >> (64 * j)) & (2**64 - 1) for j in range(4)] for x in a], dtype=np.uint64)
        B = np.array([[(x >> (64 * j)) & (2**64 - 1) for j in range(4)] for x in b], dtype=np.uint64)
        out = be.test_field_op(f, "mul", A, B)
        rinv = pow(1 << 256, -1, mod)
        got = [int(r[0]) | int(r[1]) << 64 | int(r[2]) << 128 | int(r[3]) << 192 for r in out]
        assert got == [x * y * rinv % mod for x, y in zip(a, b)]

"""GPU parity of lookup::prover::permute_expression_pair (C ABI) against the oracle's restatement."""
import random

import numpy as np
import pytest

from tests import pyref
from tests.gpu_common import be  # noqa: F401

pytestmark = pytest.mark.gpu


def _dev(torch, arr):
    return torch.from_numpy(np.ascontiguousarray(arr).view(np.int64)).cuda()


@pytest.mark.parametrize("seed,n,table_size", [(1, 1, 1), (2, 37, 5), (3, 1000, 64), (4, 5000, 5000), (5, 70000, 1 << 16), (6, (1 << 20) - 6, 1 << 16)])
def test_permute_expression_pair(be, orc, seed, n, table_size):
    import torch
    rng = random.Random(seed)
    vals = [rng.randrange(pyref.R_MOD) for _ in range(min(table_size, 2000))]
    if table_size > 2000:   # range-table style: 0 .. table_size-1, as halo2-lib's lookup tables are
        vals = list(range(table_size))
    tab_ints = (vals * (n // len(vals) + 1))[:n]
    present = sorted(set(tab_ints))
    nrng = np.random.default_rng(seed)
    pick = nrng.integers(0, len(present), n)
    inp_ints = [present[i] for i in pick]
    inp, tab = orc.fr(inp_ints), orc.fr(tab_ints)
    want_i, want_t = orc.permute_expression_pair(inp, tab)
    di, dt = _dev(torch, inp), _dev(torch, tab)
    oi = torch.empty_like(di); ot = torch.empty_like(dt)
    be.permute_expression_pair_dev(di.data_ptr(), dt.data_ptr(), n, oi.data_ptr(), ot.data_ptr())
    assert np.array_equal(oi.cpu().numpy().view(np.uint64), want_i)
    assert np.array_equal(ot.cpu().numpy().view(np.uint64), want_t)


def test_missing_value_is_constraint_failure(be, orc):
    import torch
    from spectre_b200.halo2 import BackendError
    di, dt = _dev(torch, orc.fr([1, 2, 3, 3])), _dev(torch, orc.fr([1, 2, 4, 4]))
    oi = torch.empty_like(di); ot = torch.empty_like(dt)
    with pytest.raises(BackendError, match="ConstraintSystemFailure"):
        be.permute_expression_pair_dev(di.data_ptr(), dt.data_ptr(), 4, oi.data_ptr(), ot.data_ptr())

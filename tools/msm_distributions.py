#!/usr/bin/env python3
"""MSM 2^20 device time for the scalar distributions SURVEY.md 8d lists (uniform, witness-like, edge cases)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spectre_b200 import halo2  # noqa: E402
from bench import rand_fr  # noqa: E402

R = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
MONT = (1 << 256) % R


def mont_rows(vals):
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        m = v * MONT % R
        out[i] = [(m >> (64 * j)) & (2**64 - 1) for j in range(4)]
    return out


def main():
    k = 20
    n = 1 << k
    be = halo2.Backend([0])
    pts = be.g1_fixed_base_mul(rand_fr(n, 1))
    params = halo2.ParamsKZG.from_parts(be, k, g_lagrange=pts).precompute()
    rng = np.random.default_rng(3)
    uniform = rand_fr(n, 2)
    dists = {"uniform": uniform}
    one = mont_rows([1])[0]; minus1 = mont_rows([R - 1])[0]
    dists["all_one"] = np.repeat(one[None, :], n, axis=0)
    dists["all_minus_one"] = np.repeat(minus1[None, :], n, axis=0)
    z = uniform.copy(); z[rng.random(n) < 0.7] = 0
    dists["70pct_zero_rest_uniform"] = z
    # witness-like: 70 % zero, 20 % < 2^16, 9 % < 2^104, 1 % uniform
    u = rng.random(n)
    small = mont_rows([int(x) for x in rng.integers(0, 1 << 16, 4096)])
    mid = mont_rows([int(rng.integers(0, 1 << 62)) * int(rng.integers(0, 1 << 42)) for _ in range(4096)])
    w = uniform.copy()
    w[u < 0.7] = 0
    sel = (u >= 0.7) & (u < 0.9); w[sel] = small[rng.integers(0, 4096, sel.sum())]
    sel = (u >= 0.9) & (u < 0.99); w[sel] = mid[rng.integers(0, 4096, sel.sum())]
    dists["witness_like"] = w
    b = uniform.copy(); b[:] = 0; b[rng.random(n) < 0.5] = one
    dists["boolean_column"] = b
    for name, sc in dists.items():
        t = torch.from_numpy(np.ascontiguousarray(sc).view(np.int64)).cuda()
        times = []
        for _ in range(5):
            params.commit_dev(halo2.BASIS_G_LAGRANGE, t.data_ptr(), n)
            times.append(be.last_device_ms)
        print(json.dumps({"distribution": name, "device_ms": round(float(np.median(times[1:])), 3), "stages": {k_: round(v, 3) for k_, v in be.last_msm_stage_ms.items()}}), flush=True)
    be.close()


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""MSM schedule probe on one GPU: device ms (CUDA events inside the library) and per-stage split for a few sizes while the
run-time knobs of csrc/msm.cu are varied in-process (they are read from the environment on every call):
  SPB_MSM_CHUNK            entries per accumulation chunk (default: chosen per size, 24..48)
usage: python tools/msm_probe.py [k ...]   -> JSON lines"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (rand_fr / scalar families)
from spectre_b200 import halo2  # noqa: E402


def main():
    import torch
    ks = [int(a) for a in sys.argv[1:]] or [20, 22, 23]
    be = halo2.Backend([0])
    dev = torch.device("cuda", 0)
    for k in ks:
        n = 1 << k
        hk = np.concatenate([bench.rand_fr(1 << 20, 500 + b) for b in range(max(1, n >> 20))])[:n]
        params = halo2.ParamsKZG.from_parts(be, k, g_lagrange=be.g1_fixed_base_mul(hk)).precompute()
        for dist in ("uniform", "witness_like", "all_minus_one"):
            sc = np.concatenate([bench.scalars_distribution(dist, 1 << 20, 900 + b) for b in range(max(1, n >> 20))])[:n] if dist != "all_minus_one" else bench.scalars_distribution(dist, n, 0)
            d = torch.from_numpy(sc.view(np.int64)).to(dev)
            ref = None
            for knobs in ({}, {"SPB_MSM_CHUNK": "32"}, {"SPB_MSM_CHUNK": "64"}, {"SPB_MSM_CHUNK": "96"}):
                for v in ("SPB_MSM_CHUNK",):
                    os.environ.pop(v, None)
                os.environ.update(knobs)
                ts, st = [], {}
                for it in range(6):
                    res = params.commit_dev(halo2.BASIS_G_LAGRANGE, d.data_ptr(), n)
                    if it >= 2:
                        ts.append(be.last_device_ms)
                        for a, b in be.last_msm_stage_ms.items():
                            st[a] = st.get(a, 0.0) + b / 4
                same = True if ref is None else bool(np.array_equal(ref, res))
                ref = res if ref is None else ref
                print(json.dumps({"k": k, "scalars": dist, "knobs": knobs, "device_ms": round(float(np.median(ts)), 3), "same_result": same,
                                  "stages_ms": {a: round(b, 3) for a, b in st.items()}}), flush=True)
            del d
        del params
        torch.cuda.empty_cache()
    be.close()


if __name__ == "__main__":
    main()

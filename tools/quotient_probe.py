#!/usr/bin/env python3
"""One launch each of the three quotient kernels at E = 2^22 with the sync-step shape's column counts (the workload of
bench.py's `roofline.quotient_kernels`), for ncu captures."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from spectre_b200 import halo2  # noqa: E402

be = halo2.Backend([0])
print(json.dumps(bench.quotient_roofline(torch, be, torch.device("cuda", 0), bench.measured_peaks()[0]["hbm_gbs"])))
be.close()

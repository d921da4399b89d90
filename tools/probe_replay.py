#!/usr/bin/env python3
"""Where does the k=20 replay spend stage 4 / 8b? Times the pieces separately (steady state, 3 repetitions)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spectre_b200 import halo2
from bench import rand_fr
be = halo2.Backend([0]); dev = torch.device("cuda", 0)
k = 20; n = 1 << k; usable = n - 6
pts = be.g1_fixed_base_mul(rand_fr(n, 1))
params = halo2.ParamsKZG.from_parts(be, k, g_lagrange=pts).precompute()
table = torch.zeros((n, 4), dtype=torch.int64, device=dev); table[:, 0] = torch.arange(n, device=dev) % (1 << 19)
lk_in = table[torch.randint(0, usable, (n,), device=dev)].contiguous()
pi = torch.empty_like(table); pt = torch.empty_like(table)
uni = torch.from_numpy(rand_fr(n, 2).view(np.int64)).to(dev)
def t(label, fn, reps=3):
    for i in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        print("%-40s rep %d: %8.3f ms   (last MSM stages %s)" % (label, i, (time.perf_counter() - t0) * 1e3, {a: round(b, 2) for a, b in be.last_msm_stage_ms.items()}), flush=True)
t("permute", lambda: be.permute_expression_pair_dev(lk_in.data_ptr(), table.data_ptr(), usable, pi.data_ptr(), pt.data_ptr()))
pi[usable:] = 0; pt[usable:] = 0
t("msm uniform", lambda: params.commit_dev(1, uni.data_ptr(), n))
t("msm permuted_input (sorted)", lambda: params.commit_dev(1, pi.data_ptr(), n))
t("msm permuted_table", lambda: params.commit_dev(1, pt.data_ptr(), n))
t("msm batch [pi, pt]", lambda: params.commit_batch_dev(1, [pi.data_ptr(), pt.data_ptr()], n))
be.close()

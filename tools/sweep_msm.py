#!/usr/bin/env python3
"""Sweep the MSM window width (SPB_MSM_C / SPB_MSM_C_TABLES) on the GPU box: one subprocess per setting."""
import json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for tables in (1, 0):
    for c in ([17, 18, 19, 20, 21] if tables else [15, 16, 17]):
        env = dict(os.environ); env["SPB_MSM_C_TABLES" if tables else "SPB_MSM_C"] = str(c)
        cmd = [sys.executable, os.path.join(root, "bench.py"), "--steps", "60", "--no-cpu-baseline", "--no-ntt"] + ([] if tables else ["--no-tables"])
        out = subprocess.run(cmd, env=env, capture_output=True, text=True).stdout.strip().splitlines()
        try:
            d = json.loads(out[-1])
            print(json.dumps({"tables": bool(tables), "c": c, "ms_per_step": round(d["ms_per_step"], 3), "single_ms": round(d["single_msm_device_ms"], 3),
                              "stages": {k: round(v, 3) for k, v in d["stages_ms_unpipelined"].items()}}), flush=True)
        except Exception as e:
            print("failed", tables, c, e, out[-1:] , flush=True)

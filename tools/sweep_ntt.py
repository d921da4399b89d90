#!/usr/bin/env python3
"""Sweep NTT tile / digit / thread knobs on the GPU box (one subprocess per setting)."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for digit, tile, threads in ((11, 12, 512), (11, 11, 256), (11, 11, 512), (10, 12, 512), (10, 11, 256), (10, 10, 256), (12, 12, 512), (11, 12, 256)):
    env = dict(os.environ, SPB_NTT_MAX_DIGIT=str(digit), SPB_NTT_TILE_LOG=str(tile), SPB_NTT_THREADS=str(threads))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "microbench.py"), "ntt", "k=20", "k=22", "k=24"], env=env, capture_output=True, text=True).stdout.strip().splitlines()
    print("digit=%d tile=2^%d threads=%d" % (digit, tile, threads), " | ".join(l for l in out if '"ntt"' in l).replace('"bench": "ntt", ', ''), flush=True)

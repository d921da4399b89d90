"""CPU tests of the compiled host side (include/spectre_b200_prover.hpp): host field arithmetic, Keccak and the EVM
transcript against their Python counterparts."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

from spectre_b200.transcript import EvmTranscriptWrite, keccak256
from tests import pyref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hooks():
    so = os.path.join(ROOT, "tests", "cpp", "libprover_hooks.so")
    src = os.path.join(ROOT, "tests", "cpp", "prover_hooks.cpp")
    hdr = os.path.join(ROOT, "include", "spectre_b200_prover.hpp")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", so, src])
    return ctypes.CDLL(so)


def _u(v):
    return np.array([(v >> (64 * i)) & (2**64 - 1) for i in range(4)], dtype=np.uint64)


def _i(a):
    return sum(int(a[i]) << (64 * i) for i in range(4))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize("fq", [0, 1])
def test_host_field(hooks, fq):
    m = pyref.P_MOD if fq else pyref.R_MOD
    rng = random.Random(fq)
    vals = [0, 1, 2, m - 1, m - 2, (1 << 253) % m] + [rng.randrange(m) for _ in range(200)]
    out = np.zeros(4, dtype=np.uint64)
    for a, b in zip(vals, reversed(vals)):
        hooks.ph_mul(fq, _p(_u(a)), _p(_u(b)), _p(out)); assert _i(out) == a * b % m
        hooks.ph_add(fq, _p(_u(a)), _p(_u(b)), _p(out)); assert _i(out) == (a + b) % m
        hooks.ph_sub(fq, _p(_u(a)), _p(_u(b)), _p(out)); assert _i(out) == (a - b) % m
    for a in vals[1:40]:
        hooks.ph_inv(fq, _p(_u(a)), _p(out)); assert _i(out) * a % m == 1
        hooks.ph_pow(fq, _p(_u(a)), ctypes.c_uint64(12345678901), _p(out)); assert _i(out) == pow(a, 12345678901, m)


def test_keccak_and_transcript(hooks):
    rng = random.Random(9)
    for n in (0, 1, 31, 32, 135, 136, 137, 272, 1000):
        data = bytes(rng.randrange(256) for _ in range(n))
        out = (ctypes.c_uint8 * 32)()
        hooks.ph_keccak(data, ctypes.c_size_t(n), out)
        assert bytes(out) == keccak256(data)
    digest = rng.randrange(pyref.R_MOD)
    T = EvmTranscriptWrite(digest)
    ops, vals, chal = [], [], []
    g = (1, 2)
    for step in range(60):
        op = rng.choice([0, 1, 2, 3, 3])
        ops.append(op)
        if op == 0:
            v = rng.randrange(pyref.R_MOD); vals.append(v); T.common_scalar(v)
        elif op == 1:
            v = rng.randrange(pyref.R_MOD); vals.append(v); T.write_scalar(v)
        elif op == 2:
            pt = pyref.ec_mul(g, rng.randrange(1, pyref.R_MOD)); vals += [pt[0], pt[1]]; T.write_ec_point(pt)
        else:
            chal.append(T.squeeze_challenge())
    ops_a = np.array(ops, dtype=np.int32)
    vals_a = np.stack([_u(v) for v in vals])
    chal_a = np.zeros((len(chal), 4), dtype=np.uint64)
    proof = (ctypes.c_uint8 * 8192)(); absorbed = (ctypes.c_size_t * 64)()
    hooks.ph_transcript.restype = ctypes.c_size_t
    n = hooks.ph_transcript(_p(_u(digest)), _p(ops_a), ctypes.c_size_t(len(ops)), _p(vals_a), _p(chal_a), proof, absorbed)
    assert bytes(proof[:n]) == bytes(T.proof)
    assert [_i(c) for c in chal_a] == chal
    assert list(absorbed[:len(T.absorbed)]) == T.absorbed

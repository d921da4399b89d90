#!/usr/bin/env python3
"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
Sizes are tiny so the instrumented run finishes in a minute; results are still checked against the oracle."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spectre_b200 import halo2  # noqa: E402
from oracle import oracle as orc  # noqa: E402

orc.build(); orc.lib()
be = halo2.Backend([0])
R = orc.R_MOD
root = pow(7, (R - 1) >> 28, R)
for k in (3, 9, 12, 13):
    a = orc.fr_random_chacha(1 << k, k)
    w = orc.fr([pow(root, 1 << (28 - k), R)])[0]
    assert np.array_equal(be.best_fft(a, w, k), orc.best_fft(a, w, k)), k
d = halo2.EvaluationDomain(be, 4, 9); od = orc.Domain(4, 9)
a = orc.fr_random_chacha(1 << 9, 5)
c = d.lagrange_to_coeff(a); assert np.array_equal(c, od.lagrange_to_coeff(a))
e = d.coeff_to_extended(c); assert np.array_equal(e, od.coeff_to_extended(c))
assert np.array_equal(d.extended_to_coeff(e), od.extended_to_coeff(e))
assert np.array_equal(d.divide_by_vanishing_poly(e), od.divide_by_vanishing_poly(e))
k = 9
n = 1 << k
params = halo2.ParamsKZG.setup(be, k, orc.srs_tau())
gl = orc.srs_g_lagrange(k, 0, n)
polys = [orc.fr_random_chacha(n, 100 + i) for i in range(3)]
polys[1][:] = orc.fr([1])[0]
want = [orc.g1_to_affine(orc.best_multiexp(p, gl)) for p in polys]
for p, w in zip(polys, want):
    assert np.array_equal(orc.g1_to_affine(params.commit_lagrange(p)), w)
    assert np.array_equal(orc.g1_to_affine(be.best_multiexp(p, gl)), w)
params.precompute()
for row, w in zip(params.commit_batch(halo2.BASIS_G_LAGRANGE, polys), want):
    assert np.array_equal(orc.g1_to_affine(row), w)
x = orc.fr_random_chacha(1, 7)[0]
p = orc.fr_random_chacha(1000, 8); p[::5] = 0
assert np.array_equal(be.batch_invert(p), orc.batch_invert(p))
assert np.array_equal(be.eval_polynomial(p, x), orc.eval_polynomial(p, x))
assert np.array_equal(be.kate_division(p, x), orc.kate_division(p, x))
be.grand_product(p); be.vec_mul(p, p); be.vec_axpy(p, x, p); be.vec_scale(p, x)
# round 2: device ChaCha20 Fr::random, resident plain bases, ParamsKZG::downsize (group inverse DFT), the many-query evaluation,
# a giant bucket (block-wide chain reduction) and a huge chain (grid-wide reduction), the 96-entry chunk path is size-gated and
# covered by the 2^22+ parity tests instead
import torch  # noqa: E402
t = torch.empty((777, 4), dtype=torch.int64, device="cuda")
be.fr_random_chacha_dev(0x5eed, (1 << 32) - 3, t.data_ptr(), 777)
assert np.array_equal(t.cpu().numpy().view(np.uint64), orc.fr_random_chacha(777, 0x5eed, (1 << 32) - 3))
pts = orc.g1_fixed_base_mul(orc.fr_random_chacha(300, 21))
sc = orc.fr_random_chacha(300, 22)
assert np.array_equal(orc.g1_to_affine(halo2.ParamsKZG.from_bases(be, pts).multiexp(sc)), orc.g1_to_affine(orc.best_multiexp(sc, pts)))
small = halo2.ParamsKZG.setup(be, 6, orc.srs_tau()).downsize(4)
q = orc.fr_random_chacha(16, 23)
assert np.array_equal(orc.g1_to_affine(small.commit_lagrange(q)), orc.commit_lagrange_known_tau(4, q))
dp = [torch.from_numpy(orc.fr_random_chacha(1000, 30 + i).view(np.int64)).cuda() for i in range(3)]
xs = orc.fr_random_chacha(3, 33)
ev = be.eval_polynomial_many_dev([x_.data_ptr() for x_ in dp], 1000, xs)
for i in range(3):
    assert np.array_equal(ev[i], orc.eval_polynomial(dp[i].cpu().numpy().view(np.uint64), xs[i]))
ones = np.repeat(orc.fr([1]), 1 << 13, axis=0)
gp = orc.g1_fixed_base_mul(orc.fr_random_chacha(1 << 13, 40))
assert np.array_equal(orc.g1_to_affine(be.best_multiexp(ones, gp)), orc.g1_to_affine(orc.best_multiexp(ones, gp)))
# one whole proof of the multi-column shape: graph / permutation / lookup constraint kernels, the lookup sort, the argument
# provers and the SHPLONK opener (device buffers through torch), byte-compared with the oracle-engine proof
from spectre_b200 import circuits, plonk  # noqa: E402
from spectre_b200.transcript import EvmTranscriptWrite  # noqa: E402
from tests.plonk_oracle_engine import OracleEngine, SeededRng  # noqa: E402
kp, inst = 7, [3, 1, 4]
cs = circuits.halo2lib_shape(3, 1)
fixed, adv, copies = circuits.halo2lib_witness(cs, kp, inst, lookup_bits=3, groups=12, num_gate_advice=3, num_lookup_advice=1)
proofs = []
for E in (plonk.DeviceEngine(be, halo2.ParamsKZG.setup(be, kp, orc.srs_tau()), kp, cs.degree()), OracleEngine(kp, cs.degree())):
    pk = plonk.keygen(E, cs, kp, fixed, copies)
    proofs.append(plonk.create_proof(E, pk, [inst], adv, SeededRng(1), EvmTranscriptWrite(pk.vk_digest)))
assert proofs[0] == proofs[1]
print("sanitize smoke ok (incl. a %d-byte proof), kernels launched:" % len(proofs[0]), be.kernel_launches)
be.close()

"""Proof-shaped replay: the MSM / NTT / batch-op schedule that halo2's create_proof issues for one of Spectre's
circuits (SURVEY.md 3.3 stages 3-11, sized by the table in SURVEY.md section 8), run device-resident through the
C ABI on synthetic columns.

This is NOT a proof: without a Rust host there is no circuit, witness or transcript here. It is the in-container proxy
SURVEY.md 7 ("Hard parts") prescribes for BASELINE configs 3-5 -- every primitive call a proof of that shape makes,
with the right sizes, counts and data residency -- so that "proof-generation seconds" can be reported for the GPU path
and extrapolated for the CPU path from the same primitives' measured CPU times. evaluate_h's gate arithmetic is
circuit specific; its stand-in is one fold-with-powers-of-y pass over every extended polynomial (the same HBM traffic:
each extended polynomial read once, one written; fewer multiplications per row than the real gate graph).
"""
import time

import numpy as np

from . import halo2

# (k, advice A, lookups L, permutation polys P, quotient pieces Q, max degree j, fixed+sigma+l polys F, evaluations)
SHAPES = {
    "tiny_k10": dict(k=10, A=3, L=1, P=2, Q=3, j=4, F=4, evals=5),   # test-sized
    # lightclient-circuits/config/sync_step_20.json + sha256_flex spread config (estimate)
    "sync_step_k20": dict(k=20, A=19, L=3, P=11, Q=3, j=4, F=40, evals=110),
    # config/sync_step_verifier_23.json, verified against the committed verifier contract (row 4)
    "aggregation_K23": dict(k=23, A=1, L=1, P=1, Q=4, j=5, F=10, evals=19),
    # config/committee_update_verifier_24.json (row 5)
    "aggregation_K24": dict(k=24, A=1, L=1, P=1, Q=4, j=5, F=10, evals=19),
}


def _rand_dev(torch, n, seed, dev):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    t = torch.randint(-(2**63), 2**63 - 1, (n, 4), dtype=torch.int64, device=dev, generator=g)
    t[:, 3] &= (1 << 60) - 1      # < 2^252 < r: valid Montgomery residues
    return t


def replay(be, shape_name, tau, seed=1, tables=True, verbose=False):
    """Run the schedule once (after SRS/domain setup, which a prover does at start-up) and return timings in seconds."""
    import torch
    sh = SHAPES[shape_name]
    k, A, L, P, Q, j, F, evals = (sh[x] for x in ("k", "A", "L", "P", "Q", "j", "F", "evals"))
    n = 1 << k
    dev = torch.device("cuda", be.devices[0])
    t0 = time.perf_counter()
    params = halo2.ParamsKZG.setup(be, k, tau)
    if tables:
        params.precompute()
    dom = halo2.EvaluationDomain(be, j, k)
    E = 1 << dom.extended_k
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0

    n_polys = A + 1 + P + 3 * L                       # advice + instance + permutation z + 3 per lookup
    cols = [_rand_dev(torch, n, seed + i, dev) for i in range(n_polys)]
    ext = [torch.empty((E, 4), dtype=torch.int64, device=dev) for _ in range(n_polys)]
    scratch = torch.empty((E, 4), dtype=torch.int64, device=dev)
    hq = torch.empty((n * (j - 1), 4), dtype=torch.int64, device=dev)
    y = np.array([[3, 5, 7, 11]], dtype=np.uint64)
    x = np.array([[13, 17, 19, 23]], dtype=np.uint64)
    torch.cuda.synchronize()
    GL, G = halo2.BASIS_G_LAGRANGE, halo2.BASIS_G
    stages = {}

    def timed(name, fn):
        torch.cuda.synchronize(); t = time.perf_counter(); fn(); torch.cuda.synchronize()
        stages[name] = stages.get(name, 0.0) + time.perf_counter() - t

    ptr = lambda t: t.data_ptr()
    # warm the twiddle tables / workspaces (a long-lived prover has them): one op of each kind on scratch data
    params.commit_batch_dev(GL, [ptr(cols[0])], n); dom.lagrange_to_coeff_dev(ptr(scratch[:n])); dom.coeff_to_extended_dev(ptr(cols[0]), ptr(scratch)); dom.extended_to_coeff_dev(ptr(scratch), ptr(hq))
    torch.cuda.synchronize()

    t_all = time.perf_counter()
    timed("3_advice_commit", lambda: params.commit_batch_dev(GL, [ptr(c) for c in cols[:A]], n))
    timed("4_lookup_permuted_commit", lambda: params.commit_batch_dev(GL, [ptr(cols[(A + i) % n_polys]) for i in range(2 * L)], n))

    def grand_products():
        for i in range(P + L):
            c = cols[(A + 1 + i) % n_polys]
            tmp = scratch[:n]
            tmp.copy_(c)
            be.batch_invert_dev(ptr(tmp), n)          # denominators
            be.vec_mul_dev(ptr(tmp), ptr(c), n)       # numerator / denominator
            be.grand_product_dev(ptr(tmp), n, ptr(scratch[n:2 * n]))
        params.commit_batch_dev(GL, [ptr(cols[(A + 1 + i) % n_polys]) for i in range(P + L)], n)
    timed("5_grand_products_commit", grand_products)
    timed("6_vanishing_random_commit", lambda: params.commit_batch_dev(G, [ptr(cols[0])], n))

    def to_coeff():
        for c in cols:
            dom.lagrange_to_coeff_dev(ptr(c))
    timed("7_lagrange_to_coeff", to_coeff)

    def to_extended():
        for c, e in zip(cols, ext):
            dom.coeff_to_extended_dev(ptr(c), ptr(e))
    timed("8a_coeff_to_extended", to_extended)
    # evaluate_h stand-in: fold computed + proving-key-resident extended polynomials with powers of y
    all_ext = [ptr(e) for e in ext] + [ptr(ext[i % n_polys]) for i in range(F)]
    timed("8b_quotient_fold_standin", lambda: be.lincomb_dev(all_ext, y, ptr(scratch), E))

    def vanishing():
        dom.divide_by_vanishing_poly_dev(ptr(scratch))
        dom.extended_to_coeff_dev(ptr(scratch), ptr(hq))
        params.commit_batch_dev(G, [ptr(hq[i * n:(i + 1) * n]) for i in range(Q)], n)
    timed("9_vanishing_construct_commit", vanishing)

    def evaluations():
        for i in range(evals):
            be.eval_polynomial_dev(ptr(cols[i % n_polys]), n, x)
    timed("10_evaluations", evaluations)

    def shplonk():
        opened = [ptr(c) for c in cols] + [ptr(cols[i % n_polys]) for i in range(F)]
        for s in range(4):                                    # rotation sets
            be.lincomb_dev(opened[s::4], y, ptr(scratch[:n]), n)
            be.kate_division_dev(ptr(scratch[:n]), n, x, ptr(scratch[n:2 * n]))
        params.commit_batch_dev(G, [ptr(scratch[:n]), ptr(scratch[n:2 * n])], n)
    timed("11_shplonk", shplonk)
    total = time.perf_counter() - t_all
    msm_count = A + 2 * L + (P + L) + 1 + Q + 2
    out = {"shape": shape_name, "k": k, "extended_k": dom.extended_k, "msm_count": msm_count, "ntt_n": n_polys, "ntt_ext": n_polys, "intt_ext": 1,
           "total_s": total, "setup_s": setup_s, "stages_s": {k_: round(v, 5) for k_, v in stages.items()},
           "note": "proof-shaped replay of the primitive schedule (uniform synthetic columns), not a proof"}
    if verbose:
        print(out)
    return out

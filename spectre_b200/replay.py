"""Proof-shaped replay: the MSM / NTT / batch-op schedule that halo2's create_proof issues for one of Spectre's
circuits (SURVEY.md 3.3 stages 3-11, sized by the table in SURVEY.md section 8), run device-resident through the
C ABI on synthetic columns.

This is NOT a proof: without a Rust host there is no circuit, witness or transcript here. It is the in-container proxy
SURVEY.md 7 ("Hard parts") prescribes for BASELINE configs 3-5 -- every primitive call a proof of that shape makes,
with the right sizes, counts and data residency -- so that "proof-generation seconds" can be reported for the GPU path
and extrapolated for the CPU path from the same primitives' measured CPU times. evaluate_h runs through the real
kernels (graph evaluator + permutation + lookup constraints) on a synthetic constraint system of the right shape: one
halo2-lib basic gate q*(a + b*c - d) per advice column, the permutation argument over all equality-enabled columns in
chunks of (degree - 2), and per lookup a theta-compression graph plus the five lookup terms. The SHA gate sets of the
real circuits are not modelled (their column counts are, through A).
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

    # The schedule runs twice and the second pass is reported: the first one pays the one-time costs a long-lived prover
    # (Spectre's RPC server keeps ProverState for its lifetime, prover/src/prover.rs:44-116) pays once -- lazy kernel
    # loading, workspace allocation, twiddle tables.
    first_pass_s = None
    for rep in range(2):
        stages.clear()
        t_all = time.perf_counter()
        timed("3_advice_commit", lambda: params.commit_batch_dev(GL, [ptr(c) for c in cols[:A]], n))
        # lookups: permute_expression_pair (sort + multiset walk) on a range table, then the two permuted commitments
        usable = n - 6
        table = torch.zeros((n, 4), dtype=torch.int64, device=dev); table[:, 0] = torch.arange(n, device=dev) % (1 << min(k - 1, 19))
        gsel = torch.Generator(device=dev); gsel.manual_seed(seed + 999)
        lk_in = table[torch.randint(0, usable, (n,), device=dev, generator=gsel)].contiguous()
        perm_in = torch.empty((n, 4), dtype=torch.int64, device=dev); perm_tb = torch.empty((n, 4), dtype=torch.int64, device=dev)

        def lookups():
            for _ in range(L):
                be.permute_expression_pair_dev(ptr(lk_in), ptr(table), usable, ptr(perm_in), ptr(perm_tb))
                perm_in[usable:] = 0; perm_tb[usable:] = 0     # blinding rows come from the caller's RNG
                torch.cuda.current_stream().synchronize()
                params.commit_batch_dev(GL, [ptr(perm_in), ptr(perm_tb)], n)
        timed("4_lookup_permute_and_commit", lookups)

        def grand_products():
            for i in range(P + L):
                c = cols[(A + 1 + i) % n_polys]
                tmp = scratch[:n]
                tmp.copy_(c); torch.cuda.current_stream().synchronize()   # torch's stream is not the library's
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
        # evaluate_h on the extended coset: custom gates (graph), permutation argument, lookups
        rot_scale = 1 << (dom.extended_k - k)
        ADD, SUB, MUL, HORNER = 0, 1, 2, 6
        K_INTER, K_FIXED, K_ADVICE, K_BETA, K_GAMMA, K_THETA, K_Y, K_PREV = 1, 2, 3, 6, 7, 8, 9, 10
        rotations = np.array([0, 1, 2, 3], dtype=np.int32)
        prog, t = [], 0
        for a_i in range(A):                                   # q_a * (a + b*c - d) with b,c,d at rotations 1,2,3; fold with y
            prog += [MUL, t, K_ADVICE, a_i | (1 << 16), K_ADVICE, a_i | (2 << 16)]
            prog += [ADD, t + 1, K_ADVICE, a_i, K_INTER, t]
            prog += [SUB, t + 2, K_INTER, t + 1, K_ADVICE, a_i | (3 << 16)]
            prog += [MUL, t + 3, K_FIXED, a_i % max(1, F), K_INTER, t + 2]
            prog += [HORNER | (1 << 8), t + 4, (K_PREV if a_i == 0 else K_INTER), (0 if a_i == 0 else t - 1), K_Y, 0, K_INTER, t + 3]
            t += 5
        gate_prog = np.array(prog, dtype=np.uint32)
        fixed_ext = [ptr(ext[i % n_polys]) for i in range(max(1, F))]       # proving-key cosets (resident in a real prover)
        advice_ext = [ptr(e) for e in ext[:A]]
        zero = np.zeros((1, 4), dtype=np.uint64)
        beta, gamma, theta = y, x, y
        perm_cols = [ptr(e) for e in ext[:A + 1]] + fixed_ext[:1]            # advice + instance + one constant column
        chunk = max(1, j - 2)
        n_sets = (len(perm_cols) + chunk - 1) // chunk
        z_sets = [ptr(ext[(A + 1 + i) % n_polys]) for i in range(n_sets)]
        sigma = [fixed_ext[i % len(fixed_ext)] for i in range(len(perm_cols))]
        l0, l_last, l_active = fixed_ext[0], fixed_ext[-1], fixed_ext[len(fixed_ext) // 2]
        lk_prog = np.array([HORNER | (1 << 8), 0, K_ADVICE, 0, K_THETA, 0, K_ADVICE, 0 | (1 << 16),      # compressed input
                            HORNER | (1 << 8), 1, K_FIXED, 0, K_THETA, 0, K_FIXED, 0 | (1 << 16),        # compressed table
                            ADD, 2, K_INTER, 0, K_BETA, 0, ADD, 3, K_INTER, 1, K_GAMMA, 0, MUL, 4, K_INTER, 2, K_INTER, 3], dtype=np.uint32)
        table_value = torch.empty((E, 4), dtype=torch.int64, device=dev)

        def evaluate_h():
            be.graph_evaluate_dev(gate_prog, 5 * A, 5 * A, zero, rotations, fixed_ext, advice_ext, [], zero, beta, gamma, theta, y, ptr(scratch), E, rot_scale)
            be.permutation_constraints_dev(ptr(scratch), E, rot_scale, -(6 + 1), chunk, z_sets, perm_cols, sigma, l0, l_last, l_active, beta, gamma, y, dom.extended_omega)
            for li in range(L):
                be.graph_evaluate_dev(lk_prog, 5, 5, zero, rotations, fixed_ext, [advice_ext[li % A]], [], zero, beta, gamma, theta, y, ptr(table_value), E, rot_scale)
                lp = (A + 1 + P + 3 * li) % n_polys
                be.lookup_constraints_dev(ptr(scratch), E, rot_scale, ptr(ext[lp]), ptr(ext[(lp + 1) % n_polys]), ptr(ext[(lp + 2) % n_polys]), ptr(table_value),
                                          l0, l_last, l_active, beta, gamma, y)
        timed("8b_evaluate_h", evaluate_h)

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
        if rep == 0:
            first_pass_s = total
    msm_count = A + 2 * L + (P + L) + 1 + Q + 2
    out = {"shape": shape_name, "k": k, "extended_k": dom.extended_k, "msm_count": msm_count, "ntt_n": n_polys, "ntt_ext": n_polys, "intt_ext": 1,
           "total_s": total, "first_pass_s": first_pass_s, "setup_s": setup_s, "stages_s": {k_: round(v, 5) for k_, v in stages.items()},
           "note": "proof-shaped replay of the primitive schedule (uniform synthetic columns), not a proof"}
    if verbose:
        print(out)
    return out

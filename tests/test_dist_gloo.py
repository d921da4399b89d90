"""CPU, world_size = 2, gloo: the multi-rank MSM glue (point-range sharding + all_gather of 96-byte partial sums + host
fold through the C ABI's spb_g1_sum). The per-rank partial sum, which the GPU produces on the box, is supplied here
by the CPU oracle so that the sharding / collective / fold logic is what is under test."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    from spectre_b200 import dist as spb_dist
    orc.build(); orc.lib()
    scalars = orc.fr_random_chacha(n, 0x5eed0003)
    bases = orc.g1_fixed_base_mul(orc.fr_random_chacha(n, 0x5eed0002), threads=2)
    lo, hi = spb_dist.shard_range(n, rank, world)
    # two "commitments" per rank: the batch path folds column by column
    partials = np.stack([orc.best_multiexp(scalars[lo:hi], bases[lo:hi], threads=2),
                         orc.best_multiexp(scalars[lo:hi][::-1].copy(), bases[lo:hi], threads=2)])
    folded = spb_dist.fold_partials(partials, world)
    if rank == 0:
        full0 = orc.g1_to_affine(orc.best_multiexp(scalars, bases, threads=2))
        rev = np.concatenate([scalars[a:b][::-1] for a, b in (spb_dist.shard_range(n, r, world) for r in range(world))])
        full1 = orc.g1_to_affine(orc.best_multiexp(rev, bases, threads=2))
        ok = bool(np.array_equal(orc.g1_to_affine(folded[0]), full0) and np.array_equal(orc.g1_to_affine(folded[1]), full1))
        q.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_msm_fold_world2():
    from spectre_b200 import build
    build.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 301, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_shard_ranges_cover():
    from spectre_b200 import dist as spb_dist
    for n in (0, 1, 7, 1 << 20):
        for world in (1, 2, 3, 8):
            spans = [spb_dist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))


def _scan_worker(rank, world, port, n, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    from spectre_b200 import dist as spb_dist
    orc.build(); orc.lib()
    R = orc.R_MOD
    a = orc.fr_ints(orc.fr_random_chacha(n, 0x5eed0777))
    lo, hi = spb_dist.shard_range(n, rank, world)
    local = a[lo:hi]
    z = {}

    def local_total():                      # on the box: spb_product_dev
        p = 1
        for v in local:
            p = p * v % R
        return orc.fr([p])[0]

    def seeded_scan(seed):                  # on the box: spb_grand_product_seeded_dev
        run = orc.fr_ints(seed.reshape(1, 4))[0]
        out = []
        for v in local:
            out.append(run); run = run * v % R
        z["rows"] = out
    init = orc.fr([7])[0]
    _, total = spb_dist.sharded_grand_product(local_total, seeded_scan, rank, world, init=init)
    want, run = [], 7
    for v in a:
        want.append(run); run = run * v % R
    ok = z["rows"] == want[lo:hi] and orc.fr_ints(total.reshape(1, 4))[0] == run
    q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_row_sharded_grand_product_world2():
    """SURVEY.md 8e "grand product": one all_gather of the ranks' 32-byte totals, then a seeded local scan."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_scan_worker, args=(r, 2, port, 1001, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True and q.get(timeout=5) is True

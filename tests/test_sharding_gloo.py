"""world_size-2 (and 3) gloo tests of the multi-GPU host logic on CPU: the sample partition, the broadcast of the fit
products and the single all-reduce.  The per-rank arithmetic is done by the CPU oracle here (tests may use it; the product
path uses the CUDA kernels) on this rank's slice of the counter-based normal stream -- exactly what each GPU rank computes --
so the test proves the N-rank result equals the 1-rank result for the same seed."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ppbo_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    rng = np.random.RandomState(0)
    F, D, P, B, S = 24, 3, 17, 4, 101            # S deliberately not divisible by the world size
    W, b = rng.randn(F, D) / 0.3, rng.uniform(0, 2 * np.pi, F)
    grids = rng.rand(B, P, D)
    omega_map, hess = rng.randn(F) * 0.1, -(1.0 + rng.rand(F))
    return dict(F=F, D=D, P=P, B=B, S=S, W=W, b=b, grids=grids, omega_map=omega_map, hess=hess, mustar=0.01, sf=0.2, seed=99)


def _sample_maxima(p, lo, hi):
    """what one rank computes before mu* is known: the per-sample maxima of its samples [lo, hi) of the Philox stream on every grid"""
    if hi <= lo:
        return np.zeros((p["B"], 0))
    Z = O.philox_normals(p["seed"], 0, lo * p["F"], (hi - lo) * p["F"]).reshape(hi - lo, p["F"])
    Omega = O.rff_sample_omega(p["omega_map"], p["hess"], Z)
    return np.stack([O.rff_eval_argmax(Omega, O.rff_features(p["W"], p["b"], p["grids"][bi], p["sf"]))[0] for bi in range(p["B"])])


def _reduce(mx, p):
    return np.stack([np.maximum(mx - p["mustar"], 0).sum(1), mx.sum(1), (mx ** 2).sum(1)], axis=1)


def _partial_sums(p, lo, hi):
    return _reduce(_sample_maxima(p, lo, hi), p)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ppbo_b200.iteration import Shard
        shard = Shard()
        assert (shard.rank, shard.world) == (rank, world)
        p = _problem()
        # the protocol of iteration.run_iteration: rank 1 "fits" the weights and broadcasts (omega_MAP, hess_diag); ranks >= 1
        # evaluate all samples; rank 0 "fits" the GP and broadcasts mu*; reduction + one all-reduce.  The others start from garbage.
        F = p["F"]
        pack = torch.zeros(2 * F + 1, dtype=torch.float64)
        if rank == 1:
            pack[:F] = torch.from_numpy(p["omega_map"])
            pack[F:2 * F] = torch.from_numpy(p["hess"])
        if rank == 0:
            pack[2 * F] = p["mustar"]
        shard.broadcast(pack[:2 * F], src=1)
        p["omega_map"], p["hess"] = pack[:F].numpy(), pack[F:2 * F].numpy()
        lo, hi = shard.sample_bounds(p["S"])
        assert (hi == lo) == (rank == 0)
        Zmax = _sample_maxima(p, lo, hi)                     # before mu* is known
        shard.broadcast(pack[2 * F:], src=0)
        p["mustar"] = float(pack[2 * F])
        sums = torch.from_numpy(_reduce(Zmax, p))
        shard.all_reduce_sum(sums)
        q.put((rank, lo, hi, sums.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sample_sharding_reproduces_single_rank(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p = _problem()
    full = _partial_sums(p, 0, p["S"])
    res.sort(key=lambda t: t[0])
    assert res[0][1] == 0 and res[-1][2] == p["S"]
    for (r0, lo0, hi0, _), (r1, lo1, hi1, _) in zip(res[:-1], res[1:]):
        assert hi0 == lo1                                                   # contiguous, disjoint, complete
    for _, _, _, sums in res:
        assert np.allclose(sums, full, rtol=1e-13, atol=0)                  # every rank holds the full reduction
        assert int(np.argmax(sums[:, 0])) == int(np.argmax(full[:, 0]))     # same selected direction


def test_shard_bounds_cover_everything():
    from ppbo_b200.iteration import Shard

    class Fake(Shard):
        def __init__(self, rank, world):
            self.dist, self.group, self.rank, self.world = None, None, rank, world
    for S in (0, 1, 7, 150, 32768):
        for world in (1, 2, 3, 4, 8):
            edges = [Fake(r, world).bounds(S) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == S
            assert all(a[1] == b[0] for a, b in zip(edges[:-1], edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
            # the sample partition of run_iteration: the GP-fit rank takes none when there is more than one rank
            edges = [Fake(r, world).sample_bounds(S) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == S
            assert all(a[1] == b[0] for a, b in zip(edges[:-1], edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            if world > 1:       # large S: boundaries are multiples of 128 samples (one row tile of the tensor-core kernel)
                assert sizes[0] == 0 and max(sizes[1:]) - min(sizes[1:]) <= (128 if S >= 512 * world else 1)
                assert S < 512 * world or all(lo % 128 == 0 for lo, hi in edges)
            # weighted shares (plan_shares): complete, ordered, proportional up to one row tile per boundary
            if world > 1:
                shares = [0.25] + [1.0] * (world - 1)
                edges = [Fake(r, world).sample_bounds(S, shares) for r in range(world)]
                assert edges[0][0] == 0 and edges[-1][1] == S
                assert all(a[1] == b[0] for a, b in zip(edges[:-1], edges[1:]))
                tot = sum(shares)
                for (lo, hi), w in zip(edges, shares):
                    assert abs((hi - lo) - S * w / tot) <= 256


def test_plan_shares_water_filling():
    """every rank that samples finishes at the same time; a rank whose fit ends after that time takes no samples"""
    from ppbo_b200.iteration import plan_shares
    assert plan_shares(1, 10.0, 5.0, 9.0) == [1.0]
    sh = plan_shares(2, 10.0, 6.0, 9.0)                    # rank 1 alone would end at 15 > 10: rank 0 helps from t = 10
    ends = [s_ + st for s_, st in zip(sh, (10.0, 6.0))]
    assert abs(ends[0] - ends[1]) < 1e-12 and abs(sum(sh) - 9.0) < 1e-12 and sh[0] > 0
    sh = plan_shares(8, 10.0, 6.0, 9.0)                    # seven ranks finish at 6 + 9/7 < 10: rank 0 takes none
    assert sh[0] == 0.0 and all(abs(x - 9.0 / 7) < 1e-12 for x in sh[1:])
    sh = plan_shares(4, 3.0, 2.0, 9.0)                     # steady state: short fits, everybody samples
    assert all(x > 0 for x in sh) and abs(sum(sh) - 9.0) < 1e-12 and abs((sh[1] - sh[0]) - 1.0) < 1e-12


def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32_10 (the device generator's oracle)"""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, out in kat:
        assert tuple(int(x) for x in O.philox4x32_10([ctr], key)[0]) == out
    z = O.philox_normals(5, 0, 0, 100000)
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1) < 0.02
    assert np.array_equal(O.philox_normals(5, 0, 13, 40), z[13:53])           # offset addressing

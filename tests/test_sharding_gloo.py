"""Multi-GPU host logic on CPU: frames are block-partitioned over ranks with no data-path collective; the only
cross-rank traffic is the barrier and the max-over-ranks time / a checksum reduction (world_size 2, gloo)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import homography_js_b200 as hg
from oracle import oracle as O


def test_shard_range_is_an_exact_partition():
    for n in (0, 1, 7, 64, 4096, 100000):
        for world in (1, 2, 3, 4, 8):
            got = []
            for r in range(world):
                lo, hi = hg.workloads.shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                got += list(range(lo, hi))
            assert got == list(range(n))
            sizes = [hg.workloads.shard_range(n, r, world) for r in range(world)]
            assert max(b - a for a, b in sizes) - min(b - a for a, b in sizes) <= 1


def _frame_checksum(f):
    """One frame of a small video-like stream: per-frame destiny points, oracle warp, 64-bit checksum."""
    W, H = 96, 64
    img = np.random.default_rng(1000 + f).integers(0, 256, (H, W, 4), dtype=np.uint8)
    s = np.array([0, 0, 0, H, W, 0, W, H], np.float64)
    d = s + 6.0 * np.sin(0.37 * f + np.arange(8))
    lim = [int(v) for v in O.transform_limits(O.projective_from_squares(s, d), W, H)]
    out = O.warp_inverse_geometric(img, W, H, O.projective_from_squares(d, s), lim[0], lim[1], max(lim[2], 1), max(lim[3], 1))
    return int(np.frombuffer(out.tobytes(), np.uint8).astype(np.uint64).sum()) ^ (len(out) << 20)


def _worker(rank, world, port, n_frames, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = hg.workloads.shard_range(n_frames, rank, world)
    local = torch.zeros(n_frames, dtype=torch.int64)
    for f in range(lo, hi):
        local[f] = _frame_checksum(f)
    dist.barrier()
    t = torch.tensor([float(hi - lo)], dtype=torch.float64)   # stands in for the per-rank elapsed time
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(local, op=dist.ReduceOp.SUM)               # reporting only: every frame was owned by exactly one rank
    if rank == 0:
        q.put((local.tolist(), t.item()))
    dist.destroy_process_group()


def test_two_ranks_cover_every_frame_once():
    O.build()
    n = 9
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == [_frame_checksum(f) for f in range(n)]
    assert tmax == 5.0   # ceil(9/2): the slowest rank defines the step time


def _checksum_worker(rank, world, port, sums, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = hg.workloads.shard_range(len(sums), rank, world)
    local = 0
    for f in range(lo, hi):                       # this rank's frames: sum of their 64-bit checksums mod 2^64
        local = (local + sums[f]) & 0xFFFFFFFFFFFFFFFF
    t = torch.tensor(list(hg.workloads.u64_to_halves(local)), dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)      # what bench.py's checksum_of_checksums does over NCCL
    ok = torch.tensor([1.0 if rank != 1 else 0.0], dtype=torch.float64)   # rank 1 "fails" its gate
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)     # a gate passes only when every rank passes
    if rank == 0:
        q.put((hg.workloads.halves_to_u64(int(t[0]), int(t[1])), float(ok.item())))
    dist.destroy_process_group()


def test_checksum_of_checksums_and_gates_reduce_over_ranks():
    """The streamed configs report one 64-bit sum of all per-frame checksums over all ranks and AND their gates: the two
    halves are summed separately (int64 cannot hold a sum of u64 values), a failing rank fails the job."""
    rng = np.random.default_rng(9)
    sums = [int(v) for v in rng.integers(0, 2**63, 37, dtype=np.uint64) * 2 + rng.integers(0, 2, 37, dtype=np.uint64)]
    want = sum(sums) & 0xFFFFFFFFFFFFFFFF
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_checksum_worker, args=(r, 2, port, sums, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, gate = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == want and gate == 0.0
    for v in (0, 1, 2**32 - 1, 2**32, 2**64 - 1):
        assert hg.workloads.halves_to_u64(*hg.workloads.u64_to_halves(v)) == v


def test_checksum_reference_is_position_sensitive_and_matches_its_definition():
    a = np.arange(64, dtype=np.uint8).reshape(4, 4, 4)
    px = a.reshape(-1).view("<u4").astype(object)
    want = (sum(int(p) * (((i * 2654435761) & 0xFFFFFFFF) | 1) for i, p in enumerate(px)) + 16 * 0x9E3779B97F4A7C15) % 2**64
    assert hg._abi.checksum_reference(a) == want
    b = a.copy()
    b[0, 0], b[0, 1] = a[0, 1].copy(), a[0, 0].copy()
    assert hg._abi.checksum_reference(b) != want

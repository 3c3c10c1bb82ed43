"""World-size-2 gloo tests of the multi-GPU host logic on the CPU: contiguous sharding of frames and
database rows, all-gather of per-shard top-2 results and the exact merge.  The per-shard searches are done by the
oracle here (there is no GPU); on the GPU box the same code path runs with liborb_b200 shards over NCCL (bench.py)."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total_rows, nq, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import orb_oracle
        from orb_slam_2_ros_b200 import synth
        from orb_slam_2_ros_b200._lib import TOP2_DTYPE
        from orb_slam_2_ros_b200.sharding import allgather_merge_top2, shard_range
        queries, planted, _ = synth.synth_queries(11, total_rows, nq)
        r0, r1 = shard_range(total_rows, rank, world)
        shard = synth.synth_descriptors(11, r0, r1 - r0)         # counter-based: every rank generates its own slice
        o = orb_oracle.hamming_top2(queries, shard)
        local = np.zeros(nq, TOP2_DTYPE)
        local["best_dist"], local["second_dist"] = o["best_dist"], o["second_dist"]
        local["best_idx"] = np.where(o["best_idx"] >= 0, o["best_idx"].astype(np.int64) + r0, -1)
        local["second_idx"] = np.where(o["second_idx"] >= 0, o["second_idx"].astype(np.int64) + r0, -1)
        merged = allgather_merge_top2(local)
        full = orb_oracle.hamming_top2(queries, synth.synth_descriptors(11, 0, total_rows))
        ok = (np.array_equal(merged["best_dist"], full["best_dist"]) and np.array_equal(merged["best_idx"], full["best_idx"]) and
              np.array_equal(merged["second_dist"], full["second_dist"]) and
              np.array_equal(merged["best_idx"][planted >= 0], planted[planted >= 0]))
        q.put((rank, bool(ok), int(r1 - r0)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total_rows", [4001, 2])
def test_sharded_top2_allgather_merge_gloo(oracle, total_rows):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world = 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, total_rows, 64, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert sum(n for _, _, n in res) == total_rows


def test_shard_ranges_cover_exactly():
    from orb_slam_2_ros_b200.sharding import shard_range
    for total in (0, 1, 7, 4096, 10_000_000):
        for world in (1, 2, 4, 8):
            blocks = [shard_range(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            assert all(b[1] >= b[0] for b in blocks)


def test_merge_is_order_independent_and_handles_empty_shards(oracle):
    from orb_slam_2_ros_b200 import synth, top2_merge
    from orb_slam_2_ros_b200._lib import TOP2_DTYPE
    rng = np.random.default_rng(3)
    db = synth.synth_descriptors(5, 0, 300)
    db[100] = db[7]; db[250] = db[7]                      # exact duplicates: ties on distance, lowest index must win
    qs = np.concatenate([db[[7, 100, 42]], rng.integers(0, 256, (20, 32), dtype=np.uint8)])
    cuts = [0, 0, 90, 101, 300, 300]                       # includes empty shards
    parts = np.zeros((len(cuts) - 1, len(qs)), TOP2_DTYPE)
    for s in range(len(cuts) - 1):
        o = oracle.hamming_top2(qs, db[cuts[s]:cuts[s + 1]])
        parts[s]["best_dist"], parts[s]["second_dist"] = o["best_dist"], o["second_dist"]
        parts[s]["best_idx"] = np.where(o["best_idx"] >= 0, o["best_idx"].astype(np.int64) + cuts[s], -1)
        parts[s]["second_idx"] = np.where(o["second_idx"] >= 0, o["second_idx"].astype(np.int64) + cuts[s], -1)
    full = oracle.hamming_top2(qs, db)
    for perm in ([0, 1, 2, 3, 4], [4, 2, 0, 3, 1]):
        m = top2_merge(parts[perm])
        assert np.array_equal(m["best_idx"], full["best_idx"]) and np.array_equal(m["best_dist"], full["best_dist"])
        assert np.array_equal(m["second_dist"], full["second_dist"])
    assert full["best_idx"][0] == 7 and full["best_idx"][1] == 7   # duplicates resolve to the lowest index

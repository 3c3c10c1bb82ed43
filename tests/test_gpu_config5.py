"""BASELINE config 5 at FULL size on one GPU: 2000 queries x 10 M rows, all four orb_top2 fields of all 2000 queries against the
oracle — as one shard and as 8 sequential shards merged by orb_top2_merge_device (the merge the sharded database runs after its
all-gather).  ~1 minute: the oracle side is 2 x 10^10 scalar compares spread over the host cores."""
import os
import threading

import numpy as np
import pytest

from orb_slam_2_ros_b200 import synth

pytestmark = [pytest.mark.gpu, pytest.mark.slow]
ROWS, NQ, SEED = 10_000_000, 2000, 7


def _oracle_top2(oracle, q, db):
    threads = max(1, min(32, os.cpu_count() or 1))
    parts = np.array_split(np.arange(len(q)), threads)
    out = [None] * threads

    def work(t):
        if len(parts[t]):
            out[t] = oracle.hamming_top2(q[parts[t]], db)
    th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    return np.concatenate([o for o in out if o is not None])


def test_config5_full_size_all_fields(oracle):
    import torch
    from orb_slam_2_ros_b200 import DescriptorDB
    from orb_slam_2_ros_b200._lib import TOP2_DTYPE
    from orb_slam_2_ros_b200.matcher import top2_merge_device
    from orb_slam_2_ros_b200.sharding import shard_range
    db = np.concatenate([synth.synth_descriptors(SEED, s, min(1 << 20, ROWS - s)) for s in range(0, ROWS, 1 << 20)])
    q, planted, _ = synth.synth_queries(SEED, ROWS, NQ)
    ref = _oracle_top2(oracle, q, db)
    # one shard
    one = DescriptorDB(ROWS, index_base=0)
    one.add(db)
    got = one.query_top2(q)
    for f, g in (("best_dist", "best_dist"), ("best_idx", "best_idx"), ("second_dist", "second_dist"), ("second_idx", "second_idx")):
        assert np.array_equal(got[g].astype(np.int64), ref[f].astype(np.int64)), "one shard: %s" % f
    assert np.array_equal(got["best_idx"][planted >= 0], planted[planted >= 0])
    one.close()
    # eight shards, queried one after the other, merged on the device
    parts = []
    for r in range(8):
        r0, r1 = shard_range(ROWS, r, 8)
        sh = DescriptorDB(r1 - r0, index_base=r0)
        sh.add(db[r0:r1])
        parts.append(sh.query_top2(q))
        sh.close()
    d_parts = torch.from_numpy(np.stack(parts).view(np.uint8).reshape(8, NQ, TOP2_DTYPE.itemsize)).cuda()
    d_out = torch.zeros((NQ, TOP2_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
    top2_merge_device(d_parts.data_ptr(), 8, NQ, d_out.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    merged = d_out.cpu().numpy().view(TOP2_DTYPE).reshape(NQ)
    for f in ("best_dist", "best_idx", "second_dist", "second_idx"):
        assert np.array_equal(merged[f].astype(np.int64), ref[f].astype(np.int64)), "8 shards: %s" % f

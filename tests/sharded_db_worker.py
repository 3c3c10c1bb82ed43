"""Worker of tests/test_gpu_sharded_nccl.py (one process per GPU, launched by torch.distributed.run): the library's own sharded
descriptor database — orb_db_create_sharded / orb_db_query_top2_sharded, NCCL inside liborb_b200.so — must return, on EVERY
rank, exactly what one unsharded database returns."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from orb_slam_2_ros_b200 import DescriptorDB, synth
    from orb_slam_2_ros_b200.matcher import ShardedDescriptorDB, shard_unique_id
    from orb_slam_2_ros_b200.sharding import shard_range
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")                       # only the 128-byte id travels through torch
    rows, nq = 300_001, 777
    db = synth.synth_descriptors(11, 0, rows)
    db[1000:1010] = db[5]                                 # duplicates across and inside shards: ties must resolve to the lowest global index
    db[rows - 3] = db[5]
    q, planted, _ = synth.synth_queries(11, rows, nq)
    q[0] = db[5]
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid.copy_(torch.from_numpy(shard_unique_id()))
    dist.broadcast(uid, 0)
    r0, r1 = shard_range(rows, rank, world)
    sh = ShardedDescriptorDB(r1 - r0, r0, rank, world, uid.numpy(), device=local)
    sh.add(db[r0:r1])
    got = sh.query_top2(q)
    one = DescriptorDB(rows, index_base=0, device=local)
    one.add(db)
    ref = one.query_top2(q)
    for f in ("best_dist", "best_idx", "second_dist", "second_idx"):
        assert np.array_equal(got[f], ref[f]), "rank %d: %s differs" % (rank, f)
    assert got["best_idx"][0] == 5 and got["second_idx"][0] == 1000 and got["second_dist"][0] == 0
    dist.barrier()
    if rank == 0:
        print("SHARDED_DB_OK world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

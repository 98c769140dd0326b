"""CPU, world_size 2 over gloo: the multi-GPU path is pure stream sharding -- every rank
tracks its own seeded stream with its own state, the only communication is the barrier and
the max-over-ranks of the timed region (strongsort_yolo_b200/dist.py, bench.py)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as tdist
    from helpers import FeatureBank
    from oracle import strongsort_np as ss
    from strongsort_yolo_b200 import dist, synth
    d = dist.init("gloo")
    assert dist.env_rank_world() == (rank, rank, world)
    mine = dist.streams_for_rank(rank, world, 4)
    assert mine == [rank, rank + 2]
    # each rank tracks stream `rank` independently (oracle tracker, embeddings from a FeatureBank)
    st = synth.make_stream("C1", stream_id=rank, render=False)
    trk = ss.StrongSORTOracle(None)
    bank = FeatureBank(seed=rank)
    img = np.zeros((640, 640, 3), dtype=np.uint8)
    checksum = 0.0
    for _ in range(6):
        fr = st.next_frame()
        out = trk.update(fr.dets, img, features=bank(fr.gt_ids))
        checksum += float(out[:, :5].sum()) if len(out) else 0.0
    tdist.barrier()
    slowest = dist.max_over_ranks(1.0 + rank)            # rank 1 is "slower"
    q.put((rank, checksum, trk.tracker._next_id, slowest, dist.aggregate_fps(6, slowest, world)))
    tdist.barrier()
    tdist.destroy_process_group()


def test_two_ranks_shard_by_stream():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, c0, id0, s0, f0), (r1, c1, id1, s1, f1) = res
    assert (r0, r1) == (0, 1)
    assert c0 != c1                      # different seeded streams, independent tracker state
    assert s0 == s1 == 2.0               # both ranks see the slowest rank's time
    assert f0 == f1 == 2 * 6 / 2.0       # whole-job frames / max time

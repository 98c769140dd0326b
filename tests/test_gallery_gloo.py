"""CPU, world_size 2 over gloo: the one optional collective of the path -- the all-gather of every
stream's exported tracks for the read-only cross-stream gallery (dist.gather_tracks) -- and the
match semantics (oracle/gallery_np.py) on the gathered data."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T_MAX, D = 8, 512


def _export(rank):
    """rank r exports 3 + r tracks; identity k of rank 0 reappears (noisily) as track 10 + k of rank 1."""
    rng = np.random.default_rng(100)
    base = np.maximum(rng.normal(0, 1, (6, D)), 0)
    n = 3 + rank
    feat = np.zeros((T_MAX, D), dtype=np.float32)
    ids = np.full(T_MAX, -1, dtype=np.int32)
    noise = np.random.default_rng(rank).normal(0, 0.03, (n, D))
    rows = base[:n] if rank == 0 else base[[1, 2, 4, 5]]
    f = rows + noise
    feat[:n] = (f / np.linalg.norm(f, axis=1, keepdims=True)).astype(np.float32)
    ids[:n] = np.arange(n) + (1 if rank == 0 else 10)
    return feat, ids


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as tdist
    from oracle import gallery_np
    from strongsort_yolo_b200 import dist
    dist.init("gloo")
    feat, ids = _export(rank)
    all_feat, all_ids = dist.gather_tracks(torch.from_numpy(feat), torch.from_numpy(ids))
    assert tuple(all_feat.shape) == (world, T_MAX, D) and tuple(all_ids.shape) == (world, T_MAX)
    for r in range(world):                                   # rank-major, bit-identical copies
        f_r, i_r = _export(r)
        assert np.array_equal(all_feat[r].numpy(), f_r) and np.array_equal(all_ids[r].numpy(), i_r)
    # the packed layout SharedGallery exchanges with ONE collective: [T_MAX * D float32 | T_MAX int32] per rank
    packed = torch.cat([torch.from_numpy(feat).reshape(-1).view(torch.int32), torch.from_numpy(ids)])
    allp = dist.gather_packed(packed)
    assert tuple(allp.shape) == (world, T_MAX * (D + 1))
    for r in range(world):
        f_r, i_r = _export(r)
        assert np.array_equal(allp[r, :T_MAX * D].view(torch.float32).numpy().reshape(T_MAX, D), f_r)
        assert np.array_equal(allp[r, T_MAX * D:].numpy(), i_r)
    m_rank, m_id, m_dist = gallery_np.cross_match(feat, ids, all_feat.numpy(), all_ids.numpy(), rank, 0.2)
    q.put((rank, m_rank.tolist(), m_id.tolist()))
    tdist.barrier()
    tdist.destroy_process_group()


def test_gather_and_cross_match_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29950 + (os.getpid() % 40)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict((r, (a, b)) for r, a, b in (q.get(timeout=120) for _ in procs))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # rank 0 tracks 1,2,3 = identities 0,1,2; rank 1 tracks 10..13 = identities 1,2,4,5
    assert res[0][0][:3] == [-1, 1, 1] and res[0][1][:3] == [-1, 10, 11]
    assert res[1][0][:4] == [0, 0, -1, -1] and res[1][1][:4] == [2, 3, -1, -1]

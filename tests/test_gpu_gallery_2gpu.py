"""Two GPUs, one process each over NCCL: the shared gallery's two exchanges (peer memory over NVLink through torch
symmetric memory; one NCCL all-gather of the packed export) give the same matches, and those equal the CPU restatement
(oracle/gallery_np.py) on the gathered exports.  Skipped on a box with one GPU (the driver's GPU tier); run by
tools/gpu_2gpu.sh under ``gpurun --gpus 2``."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    try:
        import torch
        import torch.distributed as tdist
        from oracle import gallery_np
        from strongsort_yolo_b200 import dist, synth
        from strongsort_yolo_b200.strong_sort import StrongSORT
        dev = torch.device("cuda", rank)
        torch.cuda.set_device(dev)
        dist.init("nccl", dev)
        # both cameras see the same scene; camera 1 starts three frames late, so its track ids differ
        st = synth.make_stream("C1")
        frames = [st.next_frame() for _ in range(16)][3 * rank:]
        trk = StrongSORT(device=str(dev), max_tracks=64, max_dets=32)
        gals = {"peer": dist.SharedGallery(trk, t_max=64, exchange="peer"),
                "nccl": dist.SharedGallery(trk, t_max=64, exchange="nccl")}
        assert gals["peer"].exchange == "peer" and gals["nccl"].exchange == "nccl"
        n_match = 0
        for f in frames[:12]:
            trk.update(f.dets, f.img)
            got = {}
            for name, g in gals.items():
                g.step()
                g.stream.synchronize()
                got[name] = [x.cpu().numpy().copy() for x in (g.m_rank, g.m_id, g.m_dist, g.ids, g.feat)]
            pr, pi, pd, ids, feat = got["peer"]
            nr, ni, nd, ids2, feat2 = got["nccl"]
            assert np.array_equal(ids, ids2) and np.array_equal(feat, feat2)
            live = ids >= 0
            assert np.array_equal(pr[live], nr[live]) and np.array_equal(pi[live], ni[live])
            assert np.array_equal(pd[live], nd[live])
            all_feat, all_ids = dist.gather_tracks(torch.from_numpy(feat).to(dev), torch.from_numpy(ids.astype(np.int32)).to(dev))
            w_rank, w_id, w_dist = gallery_np.cross_match(feat, ids, all_feat.cpu().numpy(), all_ids.cpu().numpy(), rank, 0.2)
            assert np.array_equal(pr[live], w_rank[live]) and np.array_equal(pi[live], w_id[live])
            np.testing.assert_allclose(pd[live], w_dist[live], atol=2e-6)
            n_match += int((w_rank[live] >= 0).sum())
            tdist.barrier()
        q.put((rank, "ok", n_match))
        tdist.barrier()
        tdist.destroy_process_group()
    except Exception as e:          # surface the failure in the parent
        import traceback
        q.put((rank, "error", traceback.format_exc()[-3000:]))
        raise


def test_peer_and_nccl_exchange_agree_with_the_restatement():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (tools/gpu_2gpu.sh)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + (os.getpid() % 40)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    for rank, status, info in res:
        assert status == "ok", f"rank {rank}: {info}"
    assert all(p.exitcode == 0 for p in procs)
    assert sum(info for _, _, info in res) > 10          # the two cameras do re-identify each other's tracks

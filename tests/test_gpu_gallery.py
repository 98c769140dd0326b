"""GPU: cross-stream gallery kernels (csrc/gallery.cu) against oracle/gallery_np.py through the C-ABI."""
import ctypes as C

import numpy as np
import pytest

from helpers import FeatureBank
from oracle import gallery_np, strongsort_np as ss
from strongsort_yolo_b200 import synth

pytestmark = pytest.mark.gpu


def test_export_matches_oracle_table_and_world1_step():
    from strongsort_yolo_b200 import dist
    from strongsort_yolo_b200.strong_sort import StrongSORT
    st = synth.make_stream("C1", render=False)
    bank = FeatureBank(seed=4)
    ora = ss.StrongSORTOracle(None)
    gpu = StrongSORT(max_tracks=128, max_dets=64)
    img = np.zeros((st.H, st.W, 3), dtype=np.uint8)
    for _ in range(8):
        fr = st.next_frame()
        f = bank(fr.gt_ids)
        ora.update(fr.dets, img, features=f)
        gpu.update(fr.dets, img, features=f)
    gal = dist.SharedGallery(gpu, t_max=32)
    gal.step()
    gal.stream.synchronize()
    want_f, want_i = gallery_np.export(ora.track_table())
    n = int(gal.count.item())
    assert n == len(want_i) > 0
    np.testing.assert_array_equal(gal.ids[:n].cpu().numpy(), want_i)
    assert (gal.ids[n:].cpu().numpy() == -1).all()
    np.testing.assert_allclose(gal.feat[:n].cpu().numpy(), want_f, atol=1e-5)
    assert (gal.feat[n:].cpu().numpy() == 0).all()
    assert gal.report() == []                  # a single stream has no foreign tracks


@pytest.mark.parametrize("self_rank", [0, 2])
def test_cross_match_matches_restatement(self_rank):
    import torch
    from strongsort_yolo_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(9)
    G, T, D = 4, 64, 512
    base = np.maximum(rng.normal(0, 1, (40, D)), 0)
    all_feat = np.zeros((G, T, D), dtype=np.float32)
    all_ids = np.full((G, T), -1, dtype=np.int32)
    for g in range(G):
        n = 20 + 5 * g
        who = rng.permutation(40)[:n]
        f = base[who] + rng.normal(0, 0.05, (n, D))
        all_feat[g, :n] = (f / np.linalg.norm(f, axis=1, keepdims=True)).astype(np.float32)
        all_ids[g, :n] = 1 + np.arange(n) + 100 * g
    P = lambda t: C.c_void_p(t.data_ptr())
    af, ai = torch.as_tensor(all_feat).cuda(), torch.as_tensor(all_ids).cuda()
    lf, li = af[self_rank].contiguous(), ai[self_rank].contiguous()
    m_rank = torch.zeros(T, dtype=torch.int32, device="cuda")
    m_id = torch.zeros(T, dtype=torch.int32, device="cuda")
    m_dist = torch.zeros(T, dtype=torch.float32, device="cuda")
    _lib.check(lib.ssb_gallery_cross_match(P(lf), P(li), P(af), P(ai), G, self_rank, T, D, 0.2, P(m_rank), P(m_id),
                                           P(m_dist), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    w_rank, w_id, w_dist = gallery_np.cross_match(all_feat[self_rank], all_ids[self_rank], all_feat, all_ids,
                                                  self_rank, 0.2)
    np.testing.assert_array_equal(m_rank.cpu().numpy(), w_rank)
    np.testing.assert_array_equal(m_id.cpu().numpy(), w_id)
    live = all_ids[self_rank] >= 0
    live_t = torch.as_tensor(live).cuda()
    np.testing.assert_allclose(m_dist.cpu().numpy()[live], w_dist[live], atol=2e-6)
    assert (w_rank[live] >= 0).sum() > 5            # the scenario does produce cross-stream matches
    # the packed exchange layout ([T * D float32 | T int32 ids] per rank) gives the same answer
    packed = torch.cat([af.reshape(G, -1).view(torch.int32), ai], dim=1).contiguous()
    p_rank, p_id, p_dist = torch.zeros_like(m_rank), torch.zeros_like(m_id), torch.zeros_like(m_dist)
    _lib.check(lib.ssb_gallery_cross_match_packed(P(packed), G, self_rank, T, D, 0.2, P(p_rank), P(p_id), P(p_dist),
                                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    assert torch.equal(p_rank, m_rank) and torch.equal(p_id, m_id) and torch.equal(p_dist, m_dist)
    # the collective-free kernel (foreign rows pulled through peer pointers; here all "peers" live on this GPU)
    ptrs = torch.tensor([packed[g].data_ptr() for g in range(G)], dtype=torch.int64, device="cuda")
    best = torch.zeros(T, dtype=torch.int64, device="cuda")
    q_rank, q_id, q_dist = torch.zeros_like(m_rank), torch.zeros_like(m_id), torch.zeros_like(m_dist)
    _lib.check(lib.ssb_gallery_peer_match(P(ptrs), G, self_rank, T, D, 0.2, P(best), P(q_rank), P(q_id), P(q_dist),
                                          C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    assert torch.equal(q_rank, m_rank) and torch.equal(q_id, m_id)
    assert torch.equal(q_dist[live_t], m_dist[live_t])


def test_cross_match_against_committed_golden(golden_dir):
    import os
    import torch
    from strongsort_yolo_b200 import _lib
    lib = _lib.load()
    g = np.load(os.path.join(golden_dir, "gallery_kat.npz"))
    G, T, D = g["feat"].shape
    P = lambda t: C.c_void_p(t.data_ptr())
    af, ai = torch.as_tensor(g["feat"]).cuda(), torch.as_tensor(g["ids"]).cuda()
    for r in range(G):
        m_rank = torch.zeros(T, dtype=torch.int32, device="cuda")
        m_id = torch.zeros(T, dtype=torch.int32, device="cuda")
        m_dist = torch.zeros(T, dtype=torch.float32, device="cuda")
        _lib.check(lib.ssb_gallery_cross_match(P(af[r].contiguous()), P(ai[r].contiguous()), P(af), P(ai), G, r, T, D,
                                               0.2, P(m_rank), P(m_id), P(m_dist),
                                               C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        np.testing.assert_array_equal(m_rank.cpu().numpy(), g["m_rank"][r])
        np.testing.assert_array_equal(m_id.cpu().numpy(), g["m_id"][r])

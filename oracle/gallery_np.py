"""Cross-stream gallery match, CPU restatement (TEST INFRASTRUCTURE).  The reference has no shared
gallery (its workers share nothing, /root/reference/yolo_multi_model.py:351-354): this restates the
read-only semantics DESIGN.md section 5 defines -- for every local confirmed track the nearest track
(cosine distance between unit EMA features) of any other stream, reported when <= max_dist."""
import numpy as np


def export(track_table):
    """oracle track_table() -> (feat [n,D] float32, ids [n]) of the confirmed tracks, list order."""
    keep = track_table["state"] == 2
    return track_table["feat"][keep].astype(np.float32), track_table["track_id"][keep].astype(np.int64)


def cross_match(local_feat, local_ids, all_feat, all_ids, self_rank, max_dist):
    """local [t_max,D], ids [t_max] (-1 = empty); all [G,t_max,D], ids [G,t_max].
    -> (rank [t_max], id [t_max], dist [t_max])."""
    G, t_max, D = all_feat.shape
    m_rank = np.full(t_max, -1, dtype=np.int64)
    m_id = np.full(t_max, -1, dtype=np.int64)
    m_dist = np.full(t_max, np.inf, dtype=np.float32)
    flat = all_feat.reshape(G * t_max, D).astype(np.float64)
    fid = all_ids.reshape(-1)
    foreign = (np.arange(G * t_max) // t_max != self_rank) & (fid >= 0)
    if not foreign.any():
        return m_rank, m_id, m_dist
    for i in range(t_max):
        if local_ids[i] < 0:
            continue
        d = 1.0 - flat @ local_feat[i].astype(np.float64)
        d[~foreign] = np.inf
        j = int(np.argmin(d))
        m_dist[i] = d[j]
        if d[j] <= max_dist:
            m_rank[i], m_id[i] = j // t_max, fid[j]
    return m_rank, m_id, m_dist

"""Shared helpers for the parity tests."""
import numpy as np


class FeatureBank:
    """Deterministic stand-in embeddings for tracker-only tests: one fixed
    ReLU-like vector per identity plus small per-frame noise; spurious and
    flicker detections get fresh random vectors."""

    def __init__(self, seed=0, dim=512, noise=0.05):
        self.rng = np.random.default_rng(seed)
        self.dim, self.noise = dim, noise
        self.base = {}

    def __call__(self, gt_ids):
        out = np.zeros((len(gt_ids), self.dim), dtype=np.float32)
        for i, g in enumerate(gt_ids):
            g = int(g)
            if g >= 0:
                if g not in self.base:
                    self.base[g] = np.maximum(self.rng.normal(0, 1, self.dim), 0)
                v = self.base[g] + self.noise * self.rng.normal(0, 1, self.dim)
            else:
                v = np.maximum(self.rng.normal(0, 1, self.dim), 0)
            out[i] = np.maximum(v, 0).astype(np.float32)
        return out


def assert_tables_equal(gpu, ora, rtol=1e-9, feat_tol=1e-5):
    for k in ("track_id", "state", "hits", "age", "tsu", "gallery_len"):
        np.testing.assert_array_equal(gpu[k], ora[k], err_msg=k)
    if len(ora["track_id"]):
        np.testing.assert_allclose(gpu["mean"], ora["mean"], rtol=rtol, atol=1e-9)
        np.testing.assert_allclose(gpu["cov"], ora["cov"], rtol=rtol, atol=1e-9)
        np.testing.assert_allclose(gpu["feat"], ora["feat"], rtol=0, atol=feat_tol)


def assert_rows_equal(gpu_rows, ora_rows, conf_tol=1e-6):
    assert gpu_rows.shape == ora_rows.shape, (gpu_rows.shape, ora_rows.shape)
    if gpu_rows.size == 0:
        return
    # integer boxes, ids and classes are bit-exact; conf is a float32 carried through
    np.testing.assert_array_equal(gpu_rows[:, :6], ora_rows[:, :6])
    np.testing.assert_allclose(gpu_rows[:, 6], ora_rows[:, 6], rtol=0, atol=conf_tol)

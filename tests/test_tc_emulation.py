"""CPU: the tensor-core OSBlock's host packing + index arithmetic (tests/tc_emul.py
mirrors csrc/reid_tc.cu) reproduces the fp32 torch oracle's OSBlock."""
import numpy as np
import pytest
import torch

from oracle import osnet_torch
from strongsort_yolo_b200 import weights
from tc_emul import osblock3_emul, osblock_emul

SHAPES = [(64, 32, 16), (64, 32, 64), (32, 16, 64), (32, 16, 96), (16, 8, 96), (16, 8, 128)]
NAMES = [("conv2", 0), ("conv2", 1), ("conv3", 0), ("conv3", 1), ("conv4", 0), ("conv4", 1)]


@pytest.mark.parametrize("block", [0, 1, 2, 3, 4, 5])
def test_emulated_tc_block_matches_oracle(state_dict, block):
    blob, offs = weights.pack_tc(weights.fold(state_dict))
    H, W, cin = SHAPES[block]
    rng = np.random.default_rng(block)
    x = np.maximum(rng.normal(0.5, 1.0, (2, H, W, cin)), 0).astype(np.float32)
    got = osblock_emul(block, x, blob, offs)
    model = osnet_torch.OSNet()
    sd = {k: torch.as_tensor(np.asarray(v)) for k, v in state_dict.items()}
    model.load_state_dict(sd, strict=False)
    model.eval()
    stage, idx = NAMES[block]
    with torch.no_grad():
        ref = getattr(model, stage)[idx](torch.as_tensor(x).permute(0, 3, 1, 2)).permute(0, 2, 3, 1).numpy()
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err < 2e-5, err


@pytest.mark.parametrize("block", [0, 1, 2, 3, 4, 5])
def test_emulated_tc3_block_matches_oracle(state_dict, block):
    """csrc/reid_tc3.cu (pointwise GEMM + fp32 depthwise): sections 10..15 of the blob."""
    blob, offs = weights.pack_tc(weights.fold(state_dict))
    assert len(offs) == 16
    H, W, cin = SHAPES[block]
    rng = np.random.default_rng(40 + block)
    x = np.maximum(rng.normal(0.5, 1.0, (2, H, W, cin)), 0).astype(np.float32)
    got = osblock3_emul(block, x, blob, offs)
    model = osnet_torch.OSNet()
    sd = {k: torch.as_tensor(np.asarray(v)) for k, v in state_dict.items()}
    model.load_state_dict(sd, strict=False)
    model.eval()
    stage, idx = NAMES[block]
    with torch.no_grad():
        ref = getattr(model, stage)[idx](torch.as_tensor(x).permute(0, 3, 1, 2)).permute(0, 2, 3, 1).numpy()
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err < 2e-5, err


def test_stem_space_to_depth_packing(state_dict):
    """Section 9 of the tensor-core blob (stem as 16 shifted GEMMs on the 2x2
    space-to-depth image, csrc/reid_tc.cu: stem_tc_kernel) reproduces the 7x7/2 conv."""
    import torch.nn.functional as F
    tensors = weights.fold(state_dict)
    blob, offs = weights.pack_tc(tensors)
    T = dict(tensors)
    sec = bytes(blob[int(offs[9]):])
    half = 16 * 16 * 16 * 2
    cat = np.frombuffer(sec[:2 * half], dtype=np.float16).astype(np.float64).reshape(16, 2, 32, 8)
    wt = (cat[:, :, :16] + cat[:, :, 16:]).transpose(0, 1, 3, 2).reshape(16, 16, 16)   # [tap][k][co], hi | lo rows
    bias = np.frombuffer(sec[2 * half:2 * half + 64], dtype=np.float32)
    np.testing.assert_array_equal(bias, T["stem.b"])
    rng = np.random.default_rng(0)
    R = rng.normal(0, 1, (256, 128, 3))
    w = torch.as_tensor(T["stem.w"].astype(np.float64)).permute(3, 2, 0, 1)
    ref = F.conv2d(torch.as_tensor(R).permute(2, 0, 1)[None], w, stride=2, padding=3)[0].permute(1, 2, 0).numpy()
    Rpad = np.zeros((264, 136, 3)); Rpad[3:259, 3:131] = R
    S = np.zeros((131, 67, 16))
    for dy in range(2):
        for dx in range(2):
            e = (dy * 2 + dx) * 3
            S[:, :, e:e + 3] = Rpad[dy:dy + 262:2, dx:dx + 134:2]
    out = np.zeros((128, 64, 16))
    for a in range(4):
        for b in range(4):
            out += S[a:a + 128, b:b + 64] @ wt[a * 4 + b]
    assert np.abs(out - ref).max() < 1e-5 * np.abs(ref).max()

"""CPU: the tensor-core OSBlock's host packing + index arithmetic (tests/tc_emul.py
mirrors csrc/reid_tc.cu) reproduces the fp32 torch oracle's OSBlock."""
import numpy as np
import pytest
import torch

from oracle import osnet_torch
from strongsort_yolo_b200 import weights
from tc_emul import osblock_emul

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

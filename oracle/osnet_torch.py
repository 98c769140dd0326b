"""Plain-PyTorch fp32 OSNet-x0.25 + ReID pre-processing (TEST INFRASTRUCTURE).

PARITY UNPINNED -- no ReID code or weights exist in /root/reference (SURVEY.md
section 0).  Architecture restated from Zhou et al., "Omni-Scale Feature
Learning for Person Re-Identification" (ICCV 2019) as summarised in SURVEY.md
Appendix B; module/parameter names follow the public torchreid ``osnet.py``
state_dict so that a real ``osnet_x0_25_*.pth`` checkpoint would load into both
this oracle and the CUDA path.  Pre-processing restated from SURVEY.md A.3 with
the pins written there: BGR crop (no channel swap), bilinear resize to 256x128
with half-pixel centres and no antialias (== ``F.interpolate(mode="bilinear",
align_corners=False)``), /255, ImageNet mean/std.

Self-check: ``count_macs()`` reproduces the 82.3 MMAC/crop figure of Appendix B.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REID_H, REID_W = 256, 128
PIXEL_MEAN = (0.485, 0.456, 0.406)
PIXEL_STD = (0.229, 0.224, 0.225)


class ConvLayer(nn.Module):
    def __init__(self, cin, cout, k, stride=1, padding=0):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride=stride, padding=padding, bias=False)
        self.bn = nn.BatchNorm2d(cout)

    def forward(self, x):
        return F.relu(self.bn(self.conv(x)))


class Conv1x1(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 1, bias=False)
        self.bn = nn.BatchNorm2d(cout)

    def forward(self, x):
        return F.relu(self.bn(self.conv(x)))


class Conv1x1Linear(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 1, bias=False)
        self.bn = nn.BatchNorm2d(cout)

    def forward(self, x):
        return self.bn(self.conv(x))


class LightConv3x3(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 1, bias=False)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1, bias=False, groups=cout)
        self.bn = nn.BatchNorm2d(cout)

    def forward(self, x):
        return F.relu(self.bn(self.conv2(self.conv1(x))))


class ChannelGate(nn.Module):
    def __init__(self, c, reduction=16):
        super().__init__()
        self.fc1 = nn.Conv2d(c, c // reduction, 1, bias=True)
        self.fc2 = nn.Conv2d(c // reduction, c, 1, bias=True)

    def forward(self, x):
        g = x.mean(dim=(2, 3), keepdim=True)
        g = F.relu(self.fc1(g))
        g = torch.sigmoid(self.fc2(g))
        return x * g


class OSBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        mid = cout // 4
        self.conv1 = Conv1x1(cin, mid)
        self.conv2a = LightConv3x3(mid, mid)
        self.conv2b = nn.Sequential(LightConv3x3(mid, mid), LightConv3x3(mid, mid))
        self.conv2c = nn.Sequential(*[LightConv3x3(mid, mid) for _ in range(3)])
        self.conv2d = nn.Sequential(*[LightConv3x3(mid, mid) for _ in range(4)])
        self.gate = ChannelGate(mid)
        self.conv3 = Conv1x1Linear(mid, cout)
        self.downsample = Conv1x1Linear(cin, cout) if cin != cout else None

    def forward(self, x):
        identity = x
        x1 = self.conv1(x)
        x2 = (self.gate(self.conv2a(x1)) + self.gate(self.conv2b(x1)) +
              self.gate(self.conv2c(x1)) + self.gate(self.conv2d(x1)))
        x3 = self.conv3(x2)
        if self.downsample is not None:
            identity = self.downsample(identity)
        return F.relu(x3 + identity)


class OSNet(nn.Module):
    """osnet_x0_25: channels [16, 64, 96, 128], layers [2, 2, 2], feature 512."""

    def __init__(self, channels=(16, 64, 96, 128), feature_dim=512):
        super().__init__()
        c0, c1, c2, c3 = channels
        self.conv1 = ConvLayer(3, c0, 7, stride=2, padding=3)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        self.conv2 = nn.Sequential(OSBlock(c0, c1), OSBlock(c1, c1),
                                   nn.Sequential(Conv1x1(c1, c1), nn.AvgPool2d(2, stride=2)))
        self.conv3 = nn.Sequential(OSBlock(c1, c2), OSBlock(c2, c2),
                                   nn.Sequential(Conv1x1(c2, c2), nn.AvgPool2d(2, stride=2)))
        self.conv4 = nn.Sequential(OSBlock(c2, c3), OSBlock(c3, c3))
        self.conv5 = Conv1x1(c3, c3)
        self.fc = nn.Sequential(nn.Linear(c3, feature_dim),
                                nn.BatchNorm1d(feature_dim), nn.ReLU())

    def forward(self, x):
        x = self.maxpool(self.conv1(x))
        x = self.conv5(self.conv4(self.conv3(self.conv2(x))))
        v = x.mean(dim=(2, 3))
        return self.fc(v)


def count_macs(model=None):
    """Analytic multiply-accumulate count per 3x256x128 crop."""
    model = model or OSNet()
    total = [0]

    def hook(m, inp, out):
        if isinstance(m, nn.Conv2d):
            k = m.kernel_size[0] * m.kernel_size[1]
            total[0] += out.shape[1] * out.shape[2] * out.shape[3] * k * (m.in_channels // m.groups)
        elif isinstance(m, nn.Linear):
            total[0] += m.in_features * m.out_features

    hs = [m.register_forward_hook(hook) for m in model.modules()
          if isinstance(m, (nn.Conv2d, nn.Linear))]
    model.eval()
    with torch.no_grad():
        model(torch.zeros(1, 3, REID_H, REID_W))
    for h in hs:
        h.remove()
    return total[0]


# ----------------------------------------------------------------------------
# pre-processing (A.3)
# ----------------------------------------------------------------------------
def preprocess_crops(img, boxes):
    """img: uint8 [H,W,3] (BGR, used as is); boxes: int [N,4] x1,y1,x2,y2 with
    crop = img[y1:y2, x1:x2].  Returns float32 [N,3,256,128]."""
    mean = torch.tensor(PIXEL_MEAN, dtype=torch.float32).view(1, 3, 1, 1)
    std = torch.tensor(PIXEL_STD, dtype=torch.float32).view(1, 3, 1, 1)
    out = []
    for x1, y1, x2, y2 in np.asarray(boxes).reshape(-1, 4):
        crop = np.ascontiguousarray(img[int(y1):int(y2), int(x1):int(x2)])
        if crop.shape[0] == 0 or crop.shape[1] == 0:
            # zero-area crop: an upstream crash ("the kernel must define behaviour").
            # Defined here and in stem_tc_kernel (csrc/reid_tc.cu, crop_ok): the network input is
            # all zeros AFTER normalisation (i.e. the mean colour); the detection stays in the frame.
            out.append(torch.zeros(1, 3, REID_H, REID_W))
            continue
        t = torch.from_numpy(crop).permute(2, 0, 1).unsqueeze(0).to(torch.float32)
        t = F.interpolate(t, size=(REID_H, REID_W), mode="bilinear",
                          align_corners=False)
        out.append((t / 255.0 - mean) / std)
    if not out:
        return torch.zeros(0, 3, REID_H, REID_W)
    return torch.cat(out, dim=0)


class OracleExtractor:
    """callable(img, boxes) -> float32 ndarray [N,512] using the fp32 model."""

    def __init__(self, state_dict, batch=64, threads=None):
        self.model = OSNet()
        sd = {k: torch.as_tensor(np.asarray(v)) for k, v in state_dict.items()}
        missing, unexpected = self.model.load_state_dict(sd, strict=False)
        bad = [k for k in missing if not k.endswith("num_batches_tracked")]
        if bad or unexpected:
            raise KeyError(f"state_dict mismatch: missing={bad} unexpected={unexpected}")
        self.model.eval()
        self.batch = batch
        if threads:
            torch.set_num_threads(threads)

    @torch.no_grad()
    def __call__(self, img, boxes):
        x = preprocess_crops(img, boxes)
        outs = [self.model(x[i:i + self.batch]) for i in range(0, len(x), self.batch)]
        if not outs:
            return np.zeros((0, 512), dtype=np.float32)
        return torch.cat(outs, 0).numpy().astype(np.float32)


# ----------------------------------------------------------------------------
# synthetic weights: seeded random init + BN statistics calibrated on crops
# ----------------------------------------------------------------------------
def make_synthetic_state_dict(calib_crops, seed=20240923, bias_mean=1.0,
                              res_gamma=0.3):
    """No pretrained OSNet file exists offline, so weights are Kaiming-random
    (seeded) with every BatchNorm's running statistics set to the batch
    statistics observed on ``calib_crops`` (float32 [M,3,256,128]).

    A random-weight BatchNorm network is in the chaotic regime (a 1-px crop
    shift moved the embedding by ~0.3 cosine distance), which no trained ReID
    model is.  Two choices a trained net also exhibits bring it to a realistic
    regime (same identity ~0.03, different identities ~0.6 cosine distance on
    the synthetic scenes): BN biases centred at ``bias_mean`` (ReLUs mostly in
    their linear range) and a small scale ``res_gamma`` on the last BN of every
    residual branch.  Returns {name: float32 ndarray} in torchreid naming."""
    g = torch.Generator().manual_seed(seed)
    model = OSNet()
    for name, m in model.named_modules():
        if isinstance(m, nn.Conv2d):
            fan_out = m.out_channels * m.kernel_size[0] * m.kernel_size[1] // m.groups
            with torch.no_grad():
                m.weight.normal_(0.0, (2.0 / fan_out) ** 0.5, generator=g)
                if m.bias is not None:
                    m.bias.normal_(0.0, 0.5, generator=g)
        elif isinstance(m, nn.Linear):
            with torch.no_grad():
                m.weight.normal_(0.0, (2.0 / m.in_features) ** 0.5, generator=g)
                m.bias.normal_(0.0, 0.1, generator=g)
        elif isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
            with torch.no_grad():
                m.weight.uniform_(0.6, 1.4, generator=g)
                m.bias.normal_(bias_mean, 0.25, generator=g)
                if name.endswith("conv3.bn"):
                    m.weight.mul_(res_gamma)
                    m.bias.normal_(0.0, 0.1, generator=g)
                if isinstance(m, nn.BatchNorm1d):
                    m.bias.normal_(0.0, 0.25, generator=g)
            m.momentum = 1.0  # running stats := stats of the calibration batch
    model.train()
    with torch.no_grad():
        model(calib_crops)
    model.eval()
    return {k: v.detach().cpu().numpy().astype(np.float32)
            for k, v in model.state_dict().items()
            if not k.endswith("num_batches_tracked")}

"""Generate the synthetic OSNet-x0.25 checkpoint used by tests and bench.

No pretrained ReID weights exist offline, so this script
builds a seeded random-weight OSNet whose BatchNorm statistics are calibrated
on synthetic crops (oracle/osnet_torch.make_synthetic_state_dict) and saves it
in torchreid state_dict naming as ``strongsort-yolo_b200/weights/
osnet_x0_25_synth.npz``.  A real torchreid checkpoint converted with
``np.savez(**{k: v.numpy()})`` loads through the same path.

Run from the repo root:  python tools/make_osnet_weights.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import osnet_torch, strongsort_np  # noqa: E402
from strongsort_yolo_b200 import synth  # noqa: E402


def main():
    img, dets = synth.calibration_crops()
    xywh = strongsort_np.xyxy2xywh(dets[:, :4])
    boxes = np.asarray([strongsort_np.crop_box_xyxy(b, img.shape[1], img.shape[0])
                        for b in xywh])
    crops = osnet_torch.preprocess_crops(img, boxes)
    sd = osnet_torch.make_synthetic_state_dict(crops)
    out = os.path.join(ROOT, "strongsort-yolo_b200", "weights")
    os.makedirs(out, exist_ok=True)
    path = os.path.join(out, "osnet_x0_25_synth.npz")
    np.savez(path, **sd)
    n = sum(v.size for v in sd.values())
    print(f"wrote {path}: {len(sd)} tensors, {n} parameters, "
          f"{os.path.getsize(path)/1e6:.2f} MB; MACs/crop = {osnet_torch.count_macs()}")


if __name__ == "__main__":
    main()

"""GPU: ssb_reid / ssb_update embed >= 48 crops as two halves on two streams (csrc/api.cu) -- every
crop's embedding is independent of its batch, so the result equals the per-half calls."""
import numpy as np
import pytest

from oracle import strongsort_np as ss

pytestmark = pytest.mark.gpu


def test_split_embedding_equals_unsplit():
    from strongsort_yolo_b200 import synth
    from strongsort_yolo_b200.strong_sort import StrongSORT
    st = synth.make_stream("C2")
    fr = st.next_frame()
    boxes = np.asarray([ss.crop_box_xyxy(b, 1920, 1080) for b in ss.xyxy2xywh(fr.dets[:, :4])])[:60]
    big = StrongSORT(max_tracks=64, max_dets=64)
    whole = big.extract_features(fr.img, boxes)                  # 60 crops -> split path
    a = big.extract_features(fr.img, boxes[:30])                 # 30 crops -> single stream
    b = big.extract_features(fr.img, boxes[30:])
    assert big.reid_tc_status() == 0
    np.testing.assert_allclose(whole, np.concatenate([a, b], 0), rtol=0, atol=1e-6)

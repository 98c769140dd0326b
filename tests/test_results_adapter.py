"""CPU: the Results adapter satisfies every access pattern of the reference's
process() (yolo_multi_model.py:45-162) and its label-line format (:167)."""
import numpy as np

from strongsort_yolo_b200.results import label_lines, results_from_tracks


def test_consumption_pattern_of_process():
    rows = np.array([[10, 20, 110, 220, 7, 0, 0.91], [300, 40, 360, 200, 9, 0, 0.55]])
    kp = np.random.default_rng(0).uniform(0, 100, (5, 17, 3)).astype(np.float32)
    res = results_from_tracks(rows, [3, -1], keypoints=kp)
    results = [res]
    ids = [int(bbox.id) for predictions in results if predictions is not None
           for bbox in predictions.boxes if bbox.id is not None]          # :46
    assert ids == [7, 9]
    assert res.boxes.id is not None                                        # :54
    for bbox, keypoints in zip(res.boxes, res.keypoints):                  # :58
        for keypoint in keypoints.xy.tolist():                             # :59
            assert len(keypoint) == 17 and len(keypoint[0]) == 2
    for scores, classes, bbox_coords, id_ in zip(res.boxes[0].conf, res.boxes[0].cls,
                                                  res.boxes[0].xyxy, res.boxes[0].id):   # :126
        assert int(id_) == 7 and res.names[int(classes)] == "person"       # :139
        assert [int(v) for v in bbox_coords] == [10, 20, 110, 220]
    np.testing.assert_array_equal(res.keypoints.xy[0], kp[3, :, :2])       # re-indexed to its det
    assert not res.keypoints.xy[1].any()                                   # coasting row: invalid (0,0)


def test_no_ids_frame_is_skipped_by_caller():
    res = results_from_tracks(np.zeros((0, 7)), [])
    assert res.boxes.id is None and len(list(res.boxes)) == 0


def test_label_line_format():
    rows = np.array([[10.9, 20.2, 110.7, 220.1, 7, 2, 0.91234]])
    lines = label_lines(results_from_tracks(rows, [0], names={2: "car"}))
    assert lines == ["0 2 7 0.912 10 20 110 220 -1 -1 -1 -1\n"]

"""strongsort_yolo_b200 -- B200-native (sm_100a) StrongSORT per-frame tracking path.

Drop-in for the tracker seam of bharath5673/StrongSORT-YOLO
(/root/reference/yolo_multi_model.py:41): ``StrongSORT.update(dets, img)``
backed by hand-written CUDA kernels behind a C-ABI (include/ssb.h).
There is no CPU fallback: importing the compute entry points without the built
``libssb.so`` / a CUDA device raises.
"""
__version__ = "0.1.0"

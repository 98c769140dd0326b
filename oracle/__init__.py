"""CPU oracle for the StrongSORT per-frame hot path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference snapshot (/root/reference, commit 8d118a3) ships
no StrongSORT code, no tests and no golden vectors.
Everything in this package is a CPU *restatement* of the published algorithms
(DeepSORT, StrongSORT, OSNet, scipy's rectangular LSAP) following SURVEY.md
Appendix A-C; it is self-pinned by the fixtures in tests/golden/ and, where a
third-party building block is installed here (scipy.optimize.
linear_sum_assignment, torchvision.ops.nms, torch.nn.functional), checked
against that building block.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package, and only as the checker.  The product
(strongsort_yolo_b200) never imports it and has no CPU fallback.
"""

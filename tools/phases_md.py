"""Render the in-kernel phase stamps of tools/time_stages.py (CTA 0 of every tensor-core
OSBlock launch, clock64 cycles) as a markdown table for profiles/."""
import json
import sys

d = json.load(open(sys.argv[1]))
print("| block | total cycles | phase 1 (x staging + conv1/down MMAs) | X1 drain | per 3x3 layer: wait / MMA phase / drain "
      "(median) | gate + conv3 per stream (median) | final epilogue |")
print("|---|---:|---:|---:|---|---:|---:|")
for b in range(6):
    c = d[f"osblock{b}_phase_cycles"]
    df = [c[0]] + [c[i] - c[i - 1] for i in range(1, len(c))]
    # layout: [phase1, x1drain, then per LC: start, issued, drained, (gate after last of stream)], final
    body = df[2:-1]
    waits, mmas, drains, gates = [], [], [], []
    i, lc = 0, 0
    for s in range(4):
        for k in range(s + 1):
            waits.append(body[i]); mmas.append(body[i + 1]); drains.append(body[i + 2]); i += 3
            if k == s:
                gates.append(body[i]); i += 1
    med = lambda v: sorted(v)[len(v) // 2]
    print(f"| {b} | {c[-1]} | {df[0]} | {df[1]} | {med(waits)} / {med(mmas)} / {med(drains)} | {med(gates)} | {df[-1]} |")
print()
for k, v in d.items():
    if "phase" not in k:
        print(f"* `{k}`: {v['median_us']:.1f} us (median of warm launches, CUDA events)")

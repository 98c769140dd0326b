"""Print the clock64 phase stamps of osblock3_kernel (tools/time_stages.py output) per layer."""
import json
import sys

d = json.load(open(sys.argv[1]))
for b in range(6):
    v = d[f"osblock{b}_phase_cycles"]
    print(b, "phase1", v[0], "x1drain", v[1] - v[0])
    i = 2
    prev = v[1]
    rows = []
    for s in range(4):
        for k in range(s + 1):
            t_ready, dw = v[i], v[i + 1]
            i += 2
            row = (t_ready - prev, dw - t_ready)
            prev = dw
            if k == s:
                g = v[i]
                i += 1
                row = row + (g - prev,)
                prev = g
            rows.append(row)
    print("   layers (Tdrain, dw[, gate]):", rows)
    print("   final epi", v[i] - prev, "total", v[i])

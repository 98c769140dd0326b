#!/bin/bash
# A/B knobs on the device-resident pipelined loop (bench.py --only-device): ms per frame, outputs in gpurun_out/
tag=${1:-x}
mkdir -p gpurun_out
out=gpurun_out/${tag}_exp.log
: > $out
run() { echo "== $*" >> $out; env "$@" 2>> gpurun_out/${tag}_exp.err | tail -1 >> $out; }
for c in 114688 180000; do run SSB_LSAP_SMEM_CAP=$c timeout 300 python bench.py --only-device --workload C4 --steps 30 --warmup 5; done
for c in 114688 180000; do run SSB_LSAP_SMEM_CAP=$c timeout 300 python bench.py --only-device --steps 200 --warmup 20; done
timeout 300 python tools/pipe_trace.py C4 20 > gpurun_out/${tag}_trace_c4.json 2>> gpurun_out/${tag}_exp.err
timeout 300 python tools/pipe_trace.py C2 40 > gpurun_out/${tag}_trace_c2.json 2>> gpurun_out/${tag}_exp.err
python - <<PY
import json
for c in ("c4", "c2"):
    try:
        d = json.load(open("gpurun_out/${tag}_trace_%s.json" % c))
        print(c, {k: round(v, 3) for k, v in d.items() if k.endswith("_ms") or k.endswith("_mean")})
        for r in d["rows"][6:12]: print("  ", {k: round(v, 3) for k, v in r.items()})
    except Exception as e:
        print(c, "trace:", e)
PY
cat $out

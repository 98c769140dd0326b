#!/bin/bash
# bench.py on the GPU box: the driver's default line (+ C4 block) and the reference arm; outputs in gpurun_out/
tag=${1:-x}
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
    print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "sync", round(d["e2e"]["synchronous"]["value"], 1),
          "reid_ms", round(d["roofline"]["reid_ms"], 4), "frac", round(d["roofline"]["frac"], 4))
    print("stages", {k: round(v, 1) if isinstance(v, float) else v for k, v in d["stages"].items()})
    print("cost", d["roofline_cost"])
    print("cpu", d.get("cpu_baseline"))
    print("C4", {k: v for k, v in d.get("configs", {}).get("C4", {}).items() if k != "workload"})
    print("detail", d["detail"])
except Exception as e:
    print("bench parse:", e)
PY
timeout 200 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
tail -c 400 gpurun_out/${tag}_bench_ref.json

#!/bin/bash
# short visit: LSAP + tracker parity, bench line, C4 timeline
tag=${1:-x}; mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_stages.py tests/test_gpu_tracker.py -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log; tail -2 gpurun_out/${tag}_pytest.log
bash tools/gpu_bench.sh ${tag}
timeout 200 python tools/pipe_trace.py C4 34 > gpurun_out/${tag}_trace_c4.json 2> gpurun_out/${tag}_trace.err
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_trace_c4.json"))
print("trace", d["embed_ms_mean"], d["assoc_ms_mean"], d["period_ms"], [round(r["assoc_end"] - r["assoc_start"], 2) for r in d["rows"][::4]])
PY

#!/bin/bash
# One GPU-box visit: ReID kernel parity first (bounded by `timeout`: a protocol bug must not hang the box),
# then the whole GPU suite, then the stage timings.  Outputs land in gpurun_out/ (merged back by gpurun).
tag=${1:-x}
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_reid_tc.py -x -q > gpurun_out/${tag}_reid.log 2>&1
echo "reid rc=$?" >> gpurun_out/${tag}_reid.log
tail -4 gpurun_out/${tag}_reid.log
if grep -q "rc=124" gpurun_out/${tag}_reid.log; then echo "ReID tests timed out: stopping"; exit 1; fi
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -12 gpurun_out/${tag}_pytest.log
timeout 420 python tools/time_stages.py > gpurun_out/${tag}_stages.json 2> gpurun_out/${tag}_stages.err
SSB_LSAP_SMEM_CAP=100000000 timeout 420 python tools/time_stages.py > gpurun_out/${tag}_stages_dense.json 2>> gpurun_out/${tag}_stages.err
timeout 900 python tools/build_variants.py time > gpurun_out/${tag}_variants.json 2> gpurun_out/${tag}_variants.err
cat gpurun_out/${tag}_variants.json
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_stages.json"))
    for k, v in d.items():
        if isinstance(v, dict) and "median_us" in v: print(k, round(v["median_us"], 1))
    dd = json.load(open("gpurun_out/${tag}_stages_dense.json"))
    print("lsap dense-staged:", {k: round(v["median_us"], 1) for k, v in dd.items() if k.startswith("lsap")})
except Exception as e:
    print("stages:", e)
PY

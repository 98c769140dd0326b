#!/bin/bash
# bench.py --shared-gallery on 2 GPUs with the three clock-sampling modes (A/B of the sampler's interference)
tag=${1:-x}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
p=29531
for m in off smi nvml; do
  timeout 600 $TR --master-port $p bench.py --gpus 2 --steps 60 --warmup 10 --shared-gallery --clock-sampler $m > gpurun_out/${tag}_c5_$m.json 2> gpurun_out/${tag}_c5_$m.err
  echo "$m rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_c5_$m.json").read().strip().splitlines()[-1])
    print("$m", "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "clocks", d["clocks"])
except Exception as e:
    print("$m parse:", e)
PY
  p=$((p+2))
done

#!/bin/bash
# bench.py under torchrun at N GPUs (gpurun --gpus N): the driver's launch line, default arm + reference arm
n=${1:-4}; tag=${2:-x}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_${n}gpu_devices.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29541 bench.py --gpus $n --steps 100 --warmup 10 > gpurun_out/${tag}_bench_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err
echo "bench rc=$?"; tail -c 800 gpurun_out/${tag}_bench_${n}gpu.err | grep -v "OMP_NUM_THREADS\|\*\*\*\*" | tail -5
timeout 600 $TR --master-port 29543 bench.py --impl reference --gpus $n --steps 4 --warmup 1 > gpurun_out/${tag}_bench_${n}gpu_ref.json 2>> gpurun_out/${tag}_bench_${n}gpu.err
echo "ref rc=$?"
python - <<PY
import json
for f in ("bench_${n}gpu", "bench_${n}gpu_ref"):
    try:
        d = json.loads(open("gpurun_out/${tag}_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value", d.get("value"), "n_gpus", d.get("n_gpus"), "e2e", (d.get("e2e") or {}).get("value"), "gallery", d.get("shared_gallery"))
    except Exception as e:
        print(f, "parse:", e)
PY

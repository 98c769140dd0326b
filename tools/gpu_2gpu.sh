#!/bin/bash
# Two-GPU visit (gpurun --gpus 2): the driver's N=2 launch of bench.py (+ the reference arm), the shared-gallery tests
# under NCCL, and the CLI with two sources.  Outputs in gpurun_out/
tag=${1:-x}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_2gpu_devices.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29517 bench.py --gpus 2 --steps 60 --warmup 10 > gpurun_out/${tag}_bench_2gpu.json 2> gpurun_out/${tag}_bench_2gpu.err
echo "bench2 rc=$?"; tail -c 1500 gpurun_out/${tag}_bench_2gpu.err
timeout 900 $TR --master-port 29519 bench.py --gpus 2 --steps 60 --warmup 10 --shared-gallery > gpurun_out/${tag}_bench_2gpu_c5.json 2> gpurun_out/${tag}_bench_2gpu_c5.err
echo "bench2 c5 rc=$?"; tail -c 1500 gpurun_out/${tag}_bench_2gpu_c5.err
timeout 600 $TR --master-port 29521 bench.py --impl reference --gpus 2 --steps 6 --warmup 2 > gpurun_out/${tag}_bench_2gpu_ref.json 2>> gpurun_out/${tag}_bench_2gpu.err
echo "ref2 rc=$?"
timeout 600 python -m pytest tests/test_gpu_gallery.py tests/test_gpu_gallery_2gpu.py tests/test_gpu_cli.py -q > gpurun_out/${tag}_pytest_2gpu.log 2>&1
echo "pytest2 rc=$?"; tail -5 gpurun_out/${tag}_pytest_2gpu.log
python - <<PY
import json
for f in ("bench_2gpu", "bench_2gpu_c5", "bench_2gpu_ref"):
    try:
        d = json.loads(open("gpurun_out/${tag}_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value", d.get("value"), "n_gpus", d.get("n_gpus"), "e2e", (d.get("e2e") or {}).get("value"),
              "gallery", d.get("shared_gallery") or d.get("detail", {}).get("shared_gallery"))
    except Exception as e:
        print(f, "parse:", e)
PY

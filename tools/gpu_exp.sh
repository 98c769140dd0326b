#!/bin/bash
# A/B knobs on the device-resident pipelined loop (bench.py --only-device): ms per frame, outputs in gpurun_out/
tag=${1:-x}
mkdir -p gpurun_out
out=gpurun_out/${tag}_exp.log
: > $out
run() { echo "== $*" >> $out; env "$@" 2>> gpurun_out/${tag}_exp.err | tail -1 >> $out; }
for c in 114688 180000; do run SSB_LSAP_SMEM_CAP=$c timeout 300 python bench.py --only-device --workload C4 --steps 30 --warmup 5; done
for c in 114688 180000; do run SSB_LSAP_SMEM_CAP=$c timeout 300 python bench.py --only-device --steps 200 --warmup 20; done
cat $out

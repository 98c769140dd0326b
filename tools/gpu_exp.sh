#!/bin/bash
# A/B knobs on the device-resident pipelined loop (bench.py --only-device): ms per frame, outputs in gpurun_out/
tag=${1:-x}
mkdir -p gpurun_out
out=gpurun_out/${tag}_exp.log
: > $out
run() { echo "== $*" >> $out; env "$@" 2>> gpurun_out/${tag}_exp.err | tail -1 >> $out; }
for e in 0 1; do run SSB_LSAP_EXCL=$e timeout 300 python bench.py --only-device --workload C4 --steps 30 --warmup 5; done
for e in 0 1; do run SSB_LSAP_EXCL=$e timeout 300 python bench.py --only-device --steps 200 --warmup 20; done
for s in 1 3 4; do run SSB_SPLIT=$s timeout 300 python bench.py --only-device --steps 200 --warmup 20; done
cat $out

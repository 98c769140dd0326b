#!/bin/bash
# ncu evidence of one round (1 GPU): (1) --set full of one ReID forward, caches flushed between replays (ncu default);
# (2) the same forward with --cache-control none: DRAM bytes + L2 hit rate in the warm (non-flushed) state;
# (3) --set full with source of the stage-2 identity OSBlock; (4) launch list of the pipelined bench loop.
tag=${1:-x}
mkdir -p gpurun_out
NCU="ncu --clock-control none --profile-from-start off"
timeout 900 $NCU --set full -o gpurun_out/${tag}_prof_reid -f python tools/ncu_reid.py tc > gpurun_out/${tag}_ncu1.log 2>&1
ncu -i gpurun_out/${tag}_prof_reid.ncu-rep --page raw --csv > gpurun_out/${tag}_reid_raw.csv 2>> gpurun_out/${tag}_ncu1.log
tail -2 gpurun_out/${tag}_ncu1.log
timeout 600 $NCU --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
    --csv --log-file gpurun_out/${tag}_reid_warm.csv python tools/ncu_reid.py tc > gpurun_out/${tag}_ncu2.log 2>&1
tail -1 gpurun_out/${tag}_ncu2.log
timeout 600 $NCU --set full --import-source on -k regex:osblock4 -s 1 -c 1 -o gpurun_out/${tag}_prof_k1src -f python tools/ncu_reid.py tc > gpurun_out/${tag}_ncu3.log 2>&1
ncu -i gpurun_out/${tag}_prof_k1src.ncu-rep --page source --csv > gpurun_out/${tag}_k1_source.csv 2>> gpurun_out/${tag}_ncu3.log
tail -1 gpurun_out/${tag}_ncu3.log
timeout 900 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 10 --warmup 5 --only-device > gpurun_out/${tag}_ncu4.log 2>&1
tail -1 gpurun_out/${tag}_ncu4.log
ls -la gpurun_out/${tag}_*

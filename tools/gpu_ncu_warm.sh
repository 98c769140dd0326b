#!/bin/bash
# L2 residency of the ReID forward, measured: application-replay ncu (every pass re-runs the whole program, kernels
# keep the cache state their predecessors left -- no save/restore between passes, no flush) vs the flushed capture.
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 ncu --replay-mode application --cache-control none --clock-control none --profile-from-start off \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum \
    --csv --log-file gpurun_out/${tag}_reid_apprep.csv python tools/ncu_reid.py tc > gpurun_out/${tag}_ncu5.log 2>&1
tail -2 gpurun_out/${tag}_ncu5.log

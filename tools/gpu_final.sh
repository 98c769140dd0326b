#!/bin/bash
# The round's evidence in one GPU-box visit: GPU test suite, smoke(), stage timings, bench (both arms), the two-source
# CLI, and the ncu captures (flushed --set full of one ReID forward, the same with application replay and no cache
# control, the launch list of the pipelined loop).  Outputs in gpurun_out/<tag>_*.
tag=${1:-x}
mkdir -p gpurun_out
o=gpurun_out/${tag}
timeout 1200 python -m pytest tests -m gpu -q > ${o}_pytest.log 2>&1; echo "pytest rc=$?" >> ${o}_pytest.log; tail -3 ${o}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${o}_smoke.log 2>&1; echo "smoke rc=$?" >> ${o}_smoke.log; tail -2 ${o}_smoke.log
timeout 420 python tools/time_stages.py > ${o}_stages.json 2> ${o}_stages.err
timeout 1500 python bench.py > ${o}_bench.json 2> ${o}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > ${o}_bench_ref.json 2>> ${o}_bench.err; echo "ref rc=$?"
mkdir -p /tmp/cli_run && (cd /tmp/cli_run && PYTHONPATH=$GRAFT_REPO_ROOT timeout 600 python $GRAFT_REPO_ROOT/yolo_multi_model.py --source synthetic:C1:24 synthetic:C2:14 --track --count) > ${o}_cli_2sources.log 2>&1
echo "cli rc=$?"; tail -3 ${o}_cli_2sources.log
NCU="ncu --clock-control none --profile-from-start off"
timeout 900 $NCU --set full -o ${o}_prof_reid -f python tools/ncu_reid.py tc > ${o}_ncu1.log 2>&1
ncu -i ${o}_prof_reid.ncu-rep --page raw --csv > ${o}_reid_raw.csv 2>> ${o}_ncu1.log; tail -1 ${o}_ncu1.log
timeout 900 ncu --replay-mode application --cache-control none --clock-control none --profile-from-start off \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum \
    --csv --log-file ${o}_reid_apprep.csv python tools/ncu_reid.py tc > ${o}_ncu5.log 2>&1; tail -1 ${o}_ncu5.log
timeout 900 $NCU --metrics gpu__time_duration.sum --csv --log-file ${o}_launches.csv \
    python bench.py --steps 10 --warmup 5 --only-device > ${o}_ncu4.log 2>&1; tail -1 ${o}_ncu4.log
rm -f ${o}_prof_reid.ncu-rep
python - <<PY
import json
try:
    d = json.loads(open("${o}_bench.json").read().strip().splitlines()[-1])
    print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "sync", round(d["e2e"]["synchronous"]["value"], 1),
          "reid_ms", round(d["roofline"]["reid_ms"], 4), "frac", round(d["roofline"]["frac"], 4), "assoc", round(d["stages"]["association_total"], 1))
except Exception as e:
    print("bench parse:", e)
PY
ls -la gpurun_out/ | grep ${tag}_ | awk '{print $5, $9}'

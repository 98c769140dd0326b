"""Summarise an `ncu --set full` report of one ReID forward (tools/ncu_reid.py):

  ncu -i gpurun_out/prof_reid.ncu-rep --page raw --csv > /tmp/raw.csv
  python tools/ncu_full_summary.py /tmp/raw.csv profiles/rNN_reid_ncu_full.md profiles/reid_traffic_bytes.json

Per kernel: duration, DRAM bytes read+written, DRAM throughput %, tensor-pipe
active %, achieved occupancy (warps active %), registers/thread.  The JSON holds
the summed DRAM traffic of the forward (bench.py's `roofline.traffic`).
"""
import csv
import json
import re
import sys

WANT = {
    "dur": "gpu__time_duration.sum",
    "rd": "dram__bytes_read.sum",
    "wr": "dram__bytes_write.sum",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "tensor_pct2": "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
    "warps_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
    "l2_hit": "lts__t_sector_hit_rate.pct",
    "smem_pct": "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
    "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
}
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "usecond": 1, "nsecond": 1e-3,
         "msecond": 1e3, "ms": 1e3, "second": 1e6, "s": 1e6}


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    m = re.match(r"osblock_tc_kernel<BlkCfg<([^>]*)>", name)
    if m:
        a = [s.strip() for s in m.group(1).split(",")]
        return "osblock_tc<cin %s, mid %s, cout %s, %sx%s>" % (a[0], a[1], a[3], a[4], a[5])
    m = re.match(r"osblock4_kernel<B4<([^>]*)>", name)
    if m:
        a = [s.strip() for s in m.group(1).split(",")]
        pw = len(a) > 13 and a[13] in ("1", "true", "(bool)1")
        return "osblock4<cin %s, mid %s, cout %s, %sx%s, R %s x %s bands, %s thr%s>" % (
            a[0], a[1], a[3], a[4], a[5], a[6], a[7], a[10], ", +transition" if pw else "")
    m = re.match(r"pw_tc_kernel<PwCfg<([^>]*)>", name)
    if m:
        a = [s.strip() for s in m.group(1).split(",")]
        return "pw_tc<cin %s, cout %s, %sx%s>" % (a[0], a[1], a[2], a[3])
    return re.sub(r"\(.*", "", name)


def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return None


def main():
    src, md_out, js_out = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = list(csv.reader(l for l in open(src, newline="") if l.startswith('"')))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    kn = col["Kernel Name"]
    out = []
    for r in data:
        d = {"name": short(r[kn]), "grid": r[col.get("Grid Size", kn)], "block": r[col.get("Block Size", kn)]}
        for k, m in WANT.items():
            if m in col:
                v = num(r[col[m]])
                u = units[col[m]]
                if v is not None and u in SCALE and k in ("dur", "rd", "wr"):
                    v *= SCALE[u]
                d[k] = v
        out.append(d)
    tot_b = sum((d.get("rd") or 0) + (d.get("wr") or 0) for d in out)
    tot_us = sum(d.get("dur") or 0 for d in out)
    with open(md_out, "w") as f:
        f.write("# ReID forward under `ncu --set full --clock-control none` (one C2 frame, %d launches)\n\n" % len(out))
        f.write("Per-launch times are serialised and replayed (not bench values); DRAM bytes are per launch.\n\n")
        f.write("| kernel | grid x block | us | DRAM rd MB | DRAM wr MB | DRAM %% | tensor pipe %% | SM thr %% | "
                "warps active %% | regs |\n|---|---|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for d in out:
            tp = d.get("tensor_pct")
            if tp is None:
                tp = d.get("tensor_pct2")
            f.write("| `%s` | %s x %s | %.1f | %.2f | %.2f | %s | %s | %s | %s | %s |\n" % (
                d["name"], d["grid"], d["block"], d.get("dur") or 0, (d.get("rd") or 0) / 1e6,
                (d.get("wr") or 0) / 1e6, d.get("dram_pct"), tp, d.get("sm_pct"), d.get("warps_pct"),
                int(d["regs"]) if d.get("regs") else None))
        f.write("\nsum: %.1f us, DRAM traffic %.2f MB per forward\n" % (tot_us, tot_b / 1e6))
    json.dump({"dram_bytes_per_reid_forward": tot_b, "kernels": out,
               "source": "ncu --set full --clock-control none, tools/ncu_reid.py"}, open(js_out, "w"), indent=1)
    print(open(md_out).read())


if __name__ == "__main__":
    main()

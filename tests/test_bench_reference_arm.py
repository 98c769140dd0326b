"""CPU: the reference arm of bench.py (`--impl reference`: the CPU restatement of the path on the host cores) prints
the contract's JSON line -- same metric / unit / config object as the GPU arm, `impl`, a `cpu_baseline` describing
the run and a zero-copy `e2e` block.  One timed step on a short sample, ~10 s."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["n_gpus"] == 1 and line["higher_is_better"] is True
    assert line["unit"] == "frames/s" and line["value"] > 0 and line["scaling"] == "weak" and line["data"] == "synthetic"
    sys.path.insert(0, ROOT)
    import bench
    assert line["metric"] == bench.METRIC and line["config"] == bench.base_config(1)       # identical in both arms
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "frames" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the product package never imports the oracle (the reference arm is bench.py's own, allowed, use of it)
    for root, _dirs, files in os.walk(os.path.join(ROOT, "strongsort-yolo_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(root, f)

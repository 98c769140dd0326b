"""GPU: the reference's CLI surface (`yolo_multi_model.py --source a b --track --count`,
/root/reference/yolo_multi_model.py:341-354) with TWO sources: one worker process per source through the
spawn Pool (stream i -> GPU i mod G), label files in the reference's wire format, --count from the device."""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_sources_two_workers(tmp_path):
    env = dict(os.environ, PYTHONPATH=ROOT)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "yolo_multi_model.py"), "--source",
                        "synthetic:C1:24", "synthetic:C2:14", "--track", "--count"],
                       cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    for name, frames in (("synthetic_C1_24", 24), ("synthetic_C2_14", 14)):
        path = tmp_path / "output" / f"{name}_labels.txt"
        assert path.exists(), p.stdout[-2000:]
        lines = path.read_text().strip().splitlines()
        assert len(lines) > frames                       # confirmed tracks reported over several frames
        for ln in lines[:50]:                            # reference :167 wire format
            assert re.fullmatch(r"0 \d+ \d+ \d(\.\d+)? \d+ \d+ \d+ \d+ -1 -1 -1 -1", ln), ln
        assert f"[{name}] {frames} frames" in p.stdout
        m = re.findall(r"\[%s\] count: (\{.*\})" % name, p.stdout)
        assert m, p.stdout[-2000:]
        ids = {int(ln.split()[2]) for ln in lines}
        assert sum(eval(m[-1]).values()) == len(ids)           # final overlay: every reported id counted once
    # two distinct worker processes (the Pool), each printing its own job dict (reference :245)
    assert p.stdout.count("'source': 'synthetic:") == 2

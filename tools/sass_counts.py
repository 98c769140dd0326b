"""Per-kernel SASS mnemonic counts of libssb.so (cuobjdump -sass): the instructions that prove which
hardware paths the kernels use -- UTCHMMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTCBAR (tcgen05.commit),
UBLKCP (cp.async.bulk, TMA engine, no tensor map), UTMALDG/UTMASTG (tensor-map TMA), FFMA2, DFMA, REDUX,
UCGABAR (cluster barrier), ST/LD .shared::cluster shows as ST/LD with the cluster window (not separable here).

  python tools/sass_counts.py [path/to/lib.so] > profiles/r02_sass_counts.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MNEMONICS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "UCGABAR", "FFMA2",
             "FFMA", "DFMA", "REDUX", "LDS", "STS", "LDG", "STG", "BAR"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "strongsort-yolo_b200", "libssb.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    counts, order, cur = {}, [], None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m and cur:
            op = m.group(1)
            counts[cur]["_n"] += 1
            for k in MNEMONICS:
                if op == k or (k in ("SYNCS", "BAR", "LDG", "STG", "LDS", "STS", "UCGABAR", "REDUX") and op.startswith(k)):
                    counts[cur][k] += 1
    print(f"# SASS mnemonic counts per kernel -- `{os.path.basename(lib)}` (cuobjdump -sass, sm_100a)\n")
    print("| kernel | instrs | " + " | ".join(MNEMONICS) + " |")
    print("|---|---:|" + "---:|" * len(MNEMONICS))
    tot = collections.Counter()
    for f in order:
        c = counts[f]
        name = demangle(f)
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*\)$", "", name)[:110]
        print(f"| `{name}` | {c['_n']} | " + " | ".join(str(c[k]) if c[k] else "" for k in MNEMONICS) + " |")
        tot.update(c)
    print(f"| **total** | {tot['_n']} | " + " | ".join(str(tot[k]) for k in MNEMONICS) + " |")


if __name__ == "__main__":
    main()

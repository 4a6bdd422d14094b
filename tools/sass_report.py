#!/usr/bin/env python
"""Per-kernel SASS evidence of the Blackwell-native paths in liblseg_b200.so (runs without a GPU):
counts of tcgen05 MMA (UTCHMMA, of which .2CTA and TMEM-A-operand forms), TMEM loads / stores (LDTM / STTM), TMA loads /
stores / reductions (UTMALDG / UTMASTG / UTMAREDG), packed fp32 (FFMA2 / FADD2), MUFU, legacy HMMA (must be 0).

    python tools/sass_report.py > profiles/r02_sass.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "lang-seg_b200", "liblseg_b200.so")
COLS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCHMMA tmem-A", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "FFMA2",
        "FADD2", "MUFU.EX2", "HMMA"]


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.strip().split("\n")
    return [re.sub(r"\(lseg::\w+\)|\(.*\)$", "", o).replace("void lseg::", "").replace("lseg::", "") for o in out]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    parts = re.split(r"\n\s*Function : ", sass)[1:]
    rows = []
    for p in parts:
        name, body = p.split("\n", 1)
        c = collections.Counter()
        n_inst = 0
        for line in body.split("\n"):
            m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)([^;]*);", line)
            if not m:
                continue
            n_inst += 1
            op, rest = m.group(1), m.group(2)
            base = op.split(".")[0]
            if base == "UTCHMMA":
                c["UTCHMMA"] += 1
                if ".2CTA" in op:
                    c["UTCHMMA.2CTA"] += 1
                if rest.strip().startswith("tmem["):
                    c["UTCHMMA tmem-A"] += 1
            elif base in ("LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "FFMA2", "FADD2", "HMMA"):
                c[base] += 1
            elif op.startswith("MUFU.EX2"):
                c["MUFU.EX2"] += 1
        rows.append((name.strip(), n_inst, c))
    names = demangle([r[0] for r in rows])
    print("# SASS evidence — liblseg_b200.so (cuobjdump -sass, sm_100a), round 2\n")
    print("Static instruction counts per kernel (not executed counts). `UTCHMMA` = tcgen05.mma (kind::f16), `.2CTA` = "
          "cta_group::2, `tmem-A` = A operand taken from tensor memory; `LDTM`/`STTM` = tcgen05.ld/st; `UTMALDG`/`UTMASTG`/"
          "`UTMAREDG` = cp.async.bulk.tensor load / store / reduce-add; `HMMA` (legacy mma.sync) must be 0 everywhere.\n")
    print("| kernel | instrs | " + " | ".join(COLS) + " |")
    print("|---|---:|" + "---:|" * len(COLS))
    tot = collections.Counter()
    for (_, n, c), nm in sorted(zip(rows, names), key=lambda t: t[1]):
        if not any(c[k] for k in COLS) and n < 400:
            continue
        print(f"| `{nm[:70]}` | {n} | " + " | ".join(str(c[k]) if c[k] else "" for k in COLS) + " |")
        tot.update(c)
    print("| **total** | | " + " | ".join(str(tot[k]) for k in COLS) + " |")


if __name__ == "__main__":
    main()

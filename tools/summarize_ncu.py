#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv  > profiles/r01_launches.md
  python tools/summarize_ncu.py launches gpurun_out/rn_launches.csv stem_im2col > profiles/r02_rn101_launches.md
  python tools/summarize_ncu.py full     gpurun_out/prof_gemm.ncu-rep > profiles/r01_gemm_full.md
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem",
    "launch__grid_size",
    "launch__block_size",
    "sm__cycles_elapsed.max",
]


def launches(path, marker=None):
    """marker: a kernel-name substring that starts a step (e.g. stem_im2col, patchify) — only the launches between its
    last two occurrences (ONE whole step, without the one-off weight packing / text tower of the process) are listed."""
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    if marker:
        at = [i for i, r in enumerate(rows) if marker in r["Kernel Name"]]
        if len(at) < 2:
            raise SystemExit(f"marker {marker!r} occurs {len(at)} times, need 2")
        rows = rows[at[-2]:at[-1]]
        print(f"one step = the launches between the last two `{marker}` launches of the capture\n")
    agg = collections.OrderedDict()
    for r in rows:
        name = r["Kernel Name"].split("(")[0][:70]
        t = float(r["Metric Value"].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(v[1] for v in agg.values())
    print(f"ncu launch list: {len(rows)} launches, gpu__time_duration.sum total {tot / 1e6:.3f} ms "
          f"(cold-cache, serialised: compare SHARES, not absolutes)\n")
    print("| kernel | launches | total ms | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {v[1] / tot:.3f} | {v[1] / v[0] / 1e3:.1f} |")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"ncu --set full capture: {path}\n")
    for n, r in enumerate(rows[2:]):
        print(f"### launch {n}: `{r[idx['Kernel Name']][:90]}`  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}\n")
        print("| metric | value | unit |")
        print("|---|---:|---|")
        for m in METRICS:
            if m in idx:
                print(f"| {m} | {r[idx[m]]} | {units[idx[m]]} |")
        print()


def traffic(path):
    """Per-kernel DRAM bytes (read + write) per launch from a --metrics dram__bytes_* capture."""
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = per.setdefault(r["ID"], {"name": r["Kernel Name"].split("(")[0][:40], "grid": r["Grid Size"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    agg = collections.OrderedDict()
    for d in per.values():
        a = agg.setdefault((d["name"], d["grid"]), [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0)
        a[3] += d.get("dram__bytes_write.sum", 0.0)
    print(f"DRAM traffic per launch ({len(per)} launches; dram__bytes_read.sum + dram__bytes_write.sum)\n")
    print("| kernel | grid | launches | avg us | avg read MB | avg write MB | avg traffic MB |")
    print("|---|---|---:|---:|---:|---:|---:|")
    fam = collections.OrderedDict()
    for (name, grid), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        n = a[0]
        print(f"| `{name}` | {grid} | {n} | {a[1] / n / 1e3:.1f} | {a[2] / n / 1e6:.2f} | {a[3] / n / 1e6:.2f} | "
              f"{(a[2] + a[3]) / n / 1e6:.2f} |")
        key = "gemm" if "gemm" in name else ("mhsa" if "mhsa" in name else name)
        f = fam.setdefault(key, [0, 0.0])
        f[0] += n
        f[1] += a[2] + a[3]
    print()
    for k, f in fam.items():
        print(f"family `{k}`: {f[0]} launches, mean traffic {f[1] / f[0] / 1e6:.2f} MB per launch")


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])

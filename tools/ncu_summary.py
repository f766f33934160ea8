#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/ (tracked).

    python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/rNN_launches.md [--skip N]
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/rNN_kernel.md
"""
import collections
import csv
import subprocess
import sys

FULL_KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
]


def launches(src, dst, skip=0):
    with open(src) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    rows = list(csv.DictReader(lines))[skip:]
    agg = collections.OrderedDict()
    for r in rows:
        name = r["Kernel Name"].split("(")[0]
        t = float(r["Metric Value"].replace(",", "")) / 1e3
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(v[1] for v in agg.values())
    with open(dst, "w") as out:
        out.write(f"# ncu launch list summary ({src}, first {skip} launches skipped)\n\n")
        out.write("`ncu --metrics gpu__time_duration.sum --clock-control none` — cold-cache, serialised launches: compare "
                  "SHARES, not absolutes.\n\n| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.write(f"| `{name}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / total:.1f}% |\n")
        out.write(f"\ntotal {total:.1f} us over {len(rows)} launches\n")
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as out:
        out.write(f"# ncu --set full summary ({src})\n\n")
        for r in rows[2:]:
            out.write(f"## {r[hdr.index('Kernel Name')]}  (launch id {r[hdr.index('ID')]})\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in FULL_KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    out.write(f"| {k} | {r[i]} | {units[i]} |\n")
            stalls = [(h, r[hdr.index(h)]) for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
            stalls = sorted(stalls, key=lambda kv: -float(kv[1] or 0))[:6]
            out.write("\nTop warp-stall reasons (warps stalled per issue-active cycle): " +
                      ", ".join(f"{h.split('stalled_')[1].split('_per_issue')[0]} {float(v):.2f}" for h, v in stalls) + "\n\n")
    print(open(dst).read())


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    if mode == "launches":
        launches(src, dst, int(sys.argv[5]) if len(sys.argv) > 5 and sys.argv[4] == "--skip" else 0)
    else:
        full(src, dst)

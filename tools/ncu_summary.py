#!/usr/bin/env python
"""Turn gpurun_out/ ncu captures into the small text summaries kept under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv            > profiles/rNN_launches.txt
  python tools/ncu_summary.py kernel   gpurun_out/prof_integrate.ncu-rep  > profiles/rNN_integrate_ncu.txt
  python tools/ncu_summary.py opcodes  gpurun_out/prof_integrate.ncu-rep [voxels_per_launch]

`launches` aggregates the --metrics gpu__time_duration.sum launch list per kernel (count, total, share);
`kernel` prints the metrics the roofline discussion in DESIGN.md uses from an `ncu --set full` report;
`opcodes` aggregates executed warp instructions per SASS opcode from the report's source page.
"""
import csv
import subprocess
import sys
from collections import Counter, defaultdict

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "sm__cycles_elapsed.max", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
    "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum",
]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r]
    h = next(i for i, r in enumerate(rows) if r[0] == "ID")
    hdr = rows[h]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[h + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0]
        tot[name] += float(r[mv].replace(",", ""))
        cnt[name] += 1
    total = sum(tot.values())
    print(f"# {path}: {sum(cnt.values())} launches, {total / 1e3:.1f} us of device time (ncu: cold-cache, serialised)")
    for n, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f"{n[:72]:72s} n={cnt[n]:4d} total={v / 1e3:10.1f} us avg={v / cnt[n] / 1e3:9.1f} us share={v / total:6.1%}")


def kernel(rep):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")])
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"  {m:64s} {r[i]:>16s} {units[i]}")
        rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")

        def to_bytes(v, u):
            return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        print(f"  {'dram traffic (read+write) per launch':64s} {to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr]):16.0f} byte")


def opcodes(rep, units_per_launch=None):
    rows = ncu_csv(rep, "source", ("--print-source", "sass"))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    hdr = rows[starts[0]]
    ie, src = hdr.index("Instructions Executed"), hdr.index("Source")
    end = starts[1] - 1 if len(starts) > 1 else len(rows)
    c = Counter()
    for r in rows[starts[0] + 1:end]:
        if len(r) <= ie or not r[ie].isdigit():
            continue
        tok = r[src].split()
        op = tok[1] if tok[0].startswith("@") else tok[0]
        c[op.split(".")[0]] += int(r[ie])
    total = sum(c.values())
    print(f"# first launch in {rep}: {total} warp instructions executed")
    for op, n in c.most_common(40):
        line = f"{op:10s} {n:12d} {n / total:6.1%}"
        if units_per_launch:
            line += f"  {32.0 * n / units_per_launch:7.2f} thread-instr per unit"
        print(line)


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "launches":
        launches(sys.argv[2])
    elif cmd == "kernel":
        kernel(sys.argv[2])
    elif cmd == "opcodes":
        opcodes(sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else None)
    else:
        raise SystemExit(__doc__)

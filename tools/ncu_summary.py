"""Summaries of ncu captures for profiles/ (run HERE, no GPU needed):

  python tools/ncu_summary.py launches gpurun_out/launches_r02.csv  > profiles/r02_ncu_launch_shares.txt
  python tools/ncu_summary.py full gpurun_out/prof_attn_r02.ncu-rep [more.ncu-rep ...] > profiles/r02_ncu_full.txt
  python tools/ncu_summary.py traffic FAMILY=rep.ncu-rep:ALG_BYTES ...  > profiles/r02_traffic.json

`launches`: per-kernel share of a `--metrics gpu__time_duration.sum` launch list (cold-cache, serialised: shares only).
`full`: the metrics the roofline discussion needs from `--set full` reports, one block per captured launch.
`traffic`: DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) for bench.py's roofline.traffic.
"""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__inst_executed.sum", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    units = rows[1]
    return hdr, units, rows[2:]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def cmd_full(reps):
    for rep in reps:
        hdr, units, rows = raw_rows(rep)
        ki = hdr.index("Kernel Name")
        print(f"# {rep}")
        for r in rows:
            print("---")
            print(f"{'Kernel Name':86s} {r[ki]}")
            for name in KEEP:
                if name in hdr:
                    i = hdr.index(name)
                    print(f"{name:86s} {r[i]} {units[i]}")
        print("---")


def cmd_traffic(specs):
    out = {}
    for spec in specs:
        fam, rest = spec.split("=", 1)
        rep, alg = rest.rsplit(":", 1)
        hdr, units, rows = raw_rows(rep)
        ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
        ki = hdr.index("Kernel Name")
        launches = []
        for r in rows:
            launches.append({"kernel": re.sub(r"\(.*", "", r[ki]).strip(), "dram_bytes": to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]),
                             "duration_us_under_ncu": float(r[it].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[it], 1.0)})
        tot = sum(l["dram_bytes"] for l in launches)
        out[fam] = {"dram_bytes_per_launch": tot / len(launches), "launches_captured": len(launches),
                    "algorithmic_bytes_per_launch": float(alg), "ratio": tot / len(launches) / float(alg),
                    "source": rep.split("/")[-1] + " (ncu --set full, bench shape)", "launches": launches}
    print(json.dumps(out, indent=1))


def cmd_launches(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = {}
    total = 0.0
    n = 0
    for r in rows[1:]:
        if len(r) <= mv:
            continue
        name = re.sub(r"\(.*", "", r[kn]).strip()
        name = re.sub(r"<unnamed>::", "", name)
        v = float(r[mv].replace(",", ""))
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[mu], 1e-6)
        a = agg.setdefault(name, [0.0, 0])
        a[0] += ms
        a[1] += 1
        total += ms
        n += 1
    print(f"# launches {n}  total {total:.2f} ms")
    print("share%   ms      launches  kernel")
    for name, (ms, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{100 * ms / total:6.2f} {ms:8.3f} {c:8d}  {name[:110]}")


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "full":
        cmd_full(sys.argv[2:])
    elif mode == "traffic":
        cmd_traffic(sys.argv[2:])
    else:
        cmd_launches(sys.argv[2])

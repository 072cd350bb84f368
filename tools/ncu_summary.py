"""Condenses ncu artefacts brought back in gpurun_out/ into small tracked files under profiles/.

  python tools/ncu_summary.py launches gpurun_out/r01e_launches.csv profiles/r01e_launches_summary.md
  python tools/ncu_summary.py kernel   gpurun_out/r01e_prof_walk.ncu-rep profiles/r01e_walk_kernel.json

`launches`: per-kernel count / total / mean / share of the step from the `--metrics gpu__time_duration.sum` pass
(cold-cache, serialised: compare SHARES).  `kernel`: the handful of `--set full` metrics DESIGN.md and bench.py quote
(duration, DRAM bytes, throughputs, occupancy, registers, stall reasons > 3%)."""
import collections
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        name = r[kn].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# launch list summary of {src} (ncu gpu__time_duration.sum, --clock-control none; cold-cache, serialised)\n\n")
        f.write("| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| {k} | {v[0]} | {v[1] / 1e3:.1f} | {v[1] / v[0] / 1e3:.2f} | {v[1] / tot:.3f} |\n")
    print(open(dst).read())


def kernel(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2:]
    res = {"source": src, "kernels": []}
    name_i = hdr.index("Kernel Name")
    for r in vals:
        d = {"kernel": r[name_i]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                try:
                    d[k] = {"value": float(r[i].replace(",", "")), "unit": units[i]}
                except ValueError:
                    d[k] = {"value": r[i], "unit": units[i]}
        stalls = {}
        for i, k in enumerate(hdr):
            if "warp_issue_stalled" in k and k.endswith("per_warp_active.pct"):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v > 3:
                    stalls[k.replace("smsp__warp_issue_stalled_", "").replace("_per_warp_active.pct", "")] = v
        d["stalls_pct_gt3"] = stalls
        res["kernels"].append(d)
    json.dump(res, open(dst, "w"), indent=1)
    print(json.dumps(res, indent=1)[:3000])


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])

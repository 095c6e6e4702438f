#!/usr/bin/env python
"""Turns the raw ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.
usage: summarize_ncu.py launches <launches.csv>   |   summarize_ncu.py full <report.ncu-rep>"""
import collections
import csv
import subprocess
import sys


def launches(path):
    """Per-kernel totals, split into the device-resident pass (the launch(es) with the largest grid of each kernel:
    the whole workload as one batch — what bench.py's `value` and per-kernel shares time) and the host-buffer (e2e)
    pass, whose chunked launches are smaller."""
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mn, mv, idc = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.OrderedDict()      # launch id -> [name, time_ns, grid]
    for r in data:
        if len(r) <= mv:
            continue
        # base name: template arguments dropped (the GN solver is instantiated per thread count)
        name = r[kn].split("(")[0].replace("<unnamed>::", "").replace("void ", "").split("<")[0]
        e = per.setdefault(r[idc], [name, 0.0, 0])
        if r[mn] == "gpu__time_duration.sum":
            e[1] = float(r[mv].replace(",", ""))
        elif r[mn] == "launch__grid_size":
            e[2] = int(float(r[mv].replace(",", "")))
    maxgrid = collections.defaultdict(int)
    for name, t, g in per.values():
        maxgrid[name] = max(maxgrid[name], g)
    print(f"# ncu launch list ({path}): per-launch gpu__time_duration.sum, cold-cache and serialised - compare SHARES")
    for title, pick in (("device-resident pass (largest grid per kernel = the whole workload in one batch)",
                         lambda n, g: g == maxgrid[n]),
                        ("host-buffer (e2e) pass: chunked launches", lambda n, g: g != maxgrid[n]),
                        ("all launches", lambda n, g: True)):
        agg = collections.OrderedDict()
        for name, t, g in per.values():
            if pick(name, g):
                a = agg.setdefault(name, [0, 0.0])
                a[0] += 1
                a[1] += t
        tot = sum(v[1] for v in agg.values()) or 1.0
        print(f"\n## {title}")
        print(f"{'kernel':32s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"{k:32s} {v[0]:8d} {v[1] / 1e3:12.1f} {v[1] / 1e3 / v[0]:10.1f} {v[1] / tot:7.3f}")


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "l1tex__t_bytes.sum", "smsp__inst_executed.sum"]


def full(path):
    out = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    print(f"# ncu --set full ({path}): selected raw metrics per captured launch")
    for r in rows[2:]:
        print("\n== " + r[kn].split("(")[0].replace("<unnamed>::", "").replace("void ", ""))
        extra = [h for h in hdr if h not in WANT and (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")
                                                     or h.startswith("sm__pipe_tensor") and h.endswith("pct_of_peak_sustained_active")
                                                     or h in ("smsp__issue_active.avg.pct_of_peak_sustained_active",
                                                              "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
                                                              "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
                                                              "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                                                              "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
                                                              "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
                                                              "lts__t_sector_hit_rate.pct"))]
        for w in WANT + extra:
            if w in hdr:
                i = hdr.index(w)
                if w in extra and r[i] in ("", "0", "0.0"):
                    continue
                print(f"  {w:70s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])

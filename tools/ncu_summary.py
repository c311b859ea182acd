#!/usr/bin/env python
"""Summaries of ncu output for profiles/ (run in the build container; ncu reads reports without a GPU).

  python tools/ncu_summary.py launches  <launches.csv>          per-kernel count / mean duration / share of the step
  python tools/ncu_summary.py kernel    <report.ncu-rep>        headline metrics + warp-stall samples by reason
  python tools/ncu_summary.py regions   <report.ncu-rep> name:lo:hi ...   stall samples of SASS index ranges
  python tools/ncu_summary.py traffic   <report.ncu-rep> "history=full,adapt=on" [csrc_sha16]   -> profiles/traffic.json
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "l1tex__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__block_size", "launch__grid_size",
           "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    h = rows[0]
    ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = defaultdict(list)
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1e-3)
        agg[r[ik]].append(v)
    tot = sum(sum(v) for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-62s n=%4d avg %8.1f us share %.3f" % (k[:62], len(v), sum(v) / len(v), sum(v) / tot))


def source(rep):
    rows = ncu_csv(rep, "source", ["--print-source", "sass"])
    hdr, data = rows[1], rows[2:]
    return {h: i for i, h in enumerate(hdr)}, data


def kernel(rep):
    rows = ncu_csv(rep, "raw")
    h, units, v = rows[0], rows[1], rows[2]
    print(v[h.index("Kernel Name")] if "Kernel Name" in h else "")
    for m in METRICS:
        if m in h:
            print("%-74s %s %s" % (m, v[h.index(m)], units[h.index(m)]))
    ix, data = source(rep)
    stalls = [k for k in ix if k.startswith("stall_") and "Not Issued" not in k]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    agg = sorted(((sum(int(r[ix[s]]) for r in data), s[6:]) for s in stalls), reverse=True)
    print("warp-state samples (all warps): " + ", ".join("%s %.1f%%" % (n, 100.0 * c / tot) for c, n in agg[:8]))


def regions(rep, specs):
    ix, data = source(rep)
    stalls = [k for k in ix if k.startswith("stall_") and "Not Issued" not in k]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    for spec in specs:
        name, lo, hi = spec.split(":")
        sel = data[int(lo):int(hi) + 1]
        s = sum(int(r[ix["# Samples"]]) for r in sel)
        agg = sorted(((sum(int(r[ix[k]]) for r in sel), k[6:]) for k in stalls), reverse=True)[:5]
        print("%-12s SASS %5s-%-5s %5.1f%% of samples: %s" % (name, lo, hi, 100.0 * s / tot,
              ", ".join("%s %.1f%%" % (n, 100.0 * c / tot) for c, n in agg if c)))


def traffic(rep, key, sha=None):
    """profiles/traffic.json: DRAM bytes (read + write) of the captured launch, keyed by the bench
    configuration, together with the hash of csrc/ the capture was taken with (bench.py: roofline.traffic
    is null for any other build).  `sha`: hash of the tree the capture ran on, when it differs from HEAD."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    rows = ncu_csv(rep, "raw")
    h, units, v = rows[0], rows[1], rows[2]

    def val(m):
        x = float(v[h.index(m)].replace(",", ""))
        return x * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[h.index(m)]]
    tot = int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
    path = os.path.join(root, "profiles", "traffic.json")
    sha = sha or bench.csrc_hash()
    j = json.load(open(path)) if os.path.exists(path) else {}
    if j.get("csrc_sha16") != sha:
        j = {"csrc_sha16": sha}
    j[key] = tot
    j["source"] = os.path.basename(rep)
    json.dump(j, open(path, "w"), indent=1)
    print(path, j)


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2])
    elif mode == "kernel":
        kernel(sys.argv[2])
    elif mode == "traffic":
        traffic(*sys.argv[2:5])
    else:
        regions(sys.argv[2], sys.argv[3:])

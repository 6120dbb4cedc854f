#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` output into the short text summaries kept under profiles/.

  ncu -i gpurun_out/X.ncu-rep --page raw --csv > gpurun_out/X_raw.csv
  python tools/ncu_summary.py gpurun_out/X_raw.csv [--title "..."] > profiles/rN_X.txt

With --launches it condenses a `--metrics gpu__time_duration.sum --csv` launch list instead
(per-kernel count / total / share of the listed GPU time).
"""
import argparse
import collections
import csv
import sys

KEYS = [
    # (label, metric, scale)
    ("duration [ms]", "gpu__time_duration.sum", None),
    ("SM clock [MHz]", "sm__cycles_elapsed.avg.per_second", None),
    ("grid / block", None, None),
    ("registers / thread", "launch__registers_per_thread", None),
    ("static smem / block [B]", "launch__shared_mem_per_block_static", None),
    ("local mem / thread (spill) [B]", "launch__local_mem_per_thread", None),   # may be absent
    ("theoretical occupancy [%]", "sm__maximum_warps_per_active_cycle_pct", None),
    ("achieved occupancy [%]", "sm__warps_active.avg.pct_of_peak_sustained_active", None),
    ("warp instructions executed", "smsp__inst_executed.sum", None),
    ("thread instructions executed", "smsp__thread_inst_executed.sum", None),
    ("avg active threads / warp instr", "smsp__thread_inst_executed_per_inst_executed.ratio", None),
    ("avg not-predicated-off threads / instr", "smsp__thread_inst_executed_per_inst_executed.pct", None),
    ("issue slots busy [%]", "sm__inst_issued.avg.pct_of_peak_sustained_active", None),
    ("issue slot utilisation [% of elapsed]", "sm__instruction_throughput.avg.pct_of_peak_sustained_elapsed", None),
    ("IPC (executed, per SM cycle active)", "sm__inst_executed.avg.per_cycle_active", None),
    ("eligible warps / scheduler / cycle", "smsp__warps_eligible.avg.per_cycle_active", None),
    ("issued warps / scheduler / cycle", "smsp__issue_active.avg.per_cycle_active", None),
    ("ALU pipe [% peak]", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", None),
    ("FMA pipe [% peak]", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", None),
    ("FMA-heavy pipe [% peak]", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", None),
    ("FP64 pipe [% peak]", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", None),
    ("XU pipe [% peak]", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", None),
    ("LSU pipe [% peak]", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", None),
    ("uniform pipe [% peak]", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", None),
    ("compute (SM) throughput [% peak]", "sm__throughput.avg.pct_of_peak_sustained_elapsed", None),
    ("memory throughput [% peak]", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", None),
    ("DRAM throughput [% peak]", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", None),
    ("DRAM bytes read", "dram__bytes_read.sum", None),
    ("DRAM bytes written", "dram__bytes_write.sum", None),
    ("DRAM read bandwidth", "dram__bytes_read.sum.per_second", None),
    ("L2 throughput [% peak]", "lts__throughput.avg.pct_of_peak_sustained_elapsed", None),
    ("L2 hit rate [%]", "lts__t_sector_hit_rate.pct", None),
    ("L2 sectors (total)", "lts__t_sectors.sum", None),
    ("L2 bytes (all ops)", "lts__t_bytes.sum", None),
    ("L1/TEX throughput [% peak]", "l1tex__throughput.avg.pct_of_peak_sustained_active", None),
    ("L1/TEX hit rate [%]", "l1tex__t_sector_hit_rate.pct", None),
    ("L1 global load requests", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", None),
    ("L1 global load sectors", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", None),
    ("L1->L2 read sectors (misses)", "l1tex__m_xbar2l1tex_read_sectors.sum", None),
    ("local load/store sectors", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", None),
    ("branch efficiency [%]", "smsp__sass_average_branch_targets_threads_uniform.pct", None),
]

STALLS = "smsp__average_warp_latency_issue_stalled_{}.ratio", "smsp__average_warps_issue_stalled_{}_per_issue_active.ratio"
STALL_NAMES = ["barrier", "branch_resolving", "dispatch_stall", "drain", "imc_miss", "lg_throttle", "long_scoreboard", "math_pipe_throttle",
               "membar", "mio_throttle", "misc", "no_instruction", "not_selected", "selected", "short_scoreboard", "sleeping",
               "tex_throttle", "wait"]


def fnum(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return None


def fmt(v, unit):
    x = fnum(v)
    if x is None:
        return "%s %s" % (v, unit)
    if abs(x) >= 1e6:
        return "%.4g %s" % (x, unit)
    return ("%.3f" % x).rstrip("0").rstrip(".") + " " + unit


def summarise_raw(path, title):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        out.append("# %s" % (title or path))
        out.append("kernel: %s" % r[idx["Kernel Name"]])
        out.append("grid %s  block %s  device cc %s" % (r[idx["Grid Size"]], r[idx["Block Size"]], r[idx.get("CC", 0)]))
        for label, metric, _ in KEYS:
            if metric is None or metric not in idx:
                continue
            out.append("%-42s %s" % (label, fmt(r[idx[metric]], units[idx[metric]])))
        out.append("warp stall reasons (avg warps stalled per issue-active cycle):")
        st = []
        for n in STALL_NAMES:
            for pat in STALLS:
                m = pat.format(n)
                if m in idx and fnum(r[idx[m]]) is not None:
                    st.append((fnum(r[idx[m]]), n))
                    break
        tot = sum(v for v, _ in st) or 1.0
        for v, n in sorted(st, reverse=True):
            if v / tot >= 0.005:
                out.append("  %-22s %8.3f  (%4.1f %%)" % (n, v, 100 * v / tot))
        out.append("")
    return "\n".join(out)


def summarise_launches(path, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ik]
        short = name.split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        if len(short) > 70:
            short = short[:67] + "..."
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += fnum(r[iv]) or 0.0
    tot = sum(a[1] for a in agg.values()) or 1.0
    out = ["# %s" % (title or path), "# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes",
           "%-72s %6s %12s %10s %7s" % ("kernel", "count", "total [us]", "avg [us]", "share")]
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%-72s %6d %12.1f %10.1f %6.1f%%" % (k, c, t / 1e3, t / 1e3 / c, 100 * t / tot))
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--title", default="")
    ap.add_argument("--launches", action="store_true")
    a = ap.parse_args()
    sys.stdout.write(summarise_launches(a.csv, a.title) if a.launches else summarise_raw(a.csv, a.title))

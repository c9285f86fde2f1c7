"""Turn ncu outputs into the small summaries committed under profiles/.

    python scripts/summarise_ncu.py launches <launches.csv> <out.md>
    python scripts/summarise_ncu.py full <report.ncu-rep> <out.json> [kernel-substring]
"""
import collections
import csv
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0]
        ns = float(r[-1].replace(",", ""))
        a = agg.setdefault(name, [0, 0.])
        a[0] += 1
        a[1] += ns
    total = sum(v[1] for v in agg.values())
    with open(out, "w") as fh:
        fh.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for name, (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write("| `%s` | %d | %.3f | %.1f%% |\n" % (name, cnt, ns / 1e6, 100 * ns / total))
        fh.write("\n(total %.3f ms over %d launches; per-launch times from `ncu --metrics gpu__time_duration.sum "
                 "--clock-control none`, cold-cache and serialised)\n" % (total / 1e6, len(rows)))


def full(rep, out, pattern=""):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    result = []
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        if pattern and pattern not in d.get("Kernel Name", ""):
            continue
        rec = {"kernel": d.get("Kernel Name"), "grid": d.get("Grid Size"), "block": d.get("Block Size")}
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP:
                try:
                    rec[h + (" [%s]" % u if u else "")] = float(v.replace(",", ""))
                except ValueError:
                    rec[h] = v
        result.append(rec)
    first = result[0]
    rd = [v for k, v in first.items() if k.startswith("dram__bytes_read.sum [")][0]
    wr = [v for k, v in first.items() if k.startswith("dram__bytes_write.sum [")][0]
    unit = [k for k in first if k.startswith("dram__bytes_read.sum [")][0].split("[")[1].rstrip("]")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    unit_w = [k for k in first if k.startswith("dram__bytes_write.sum [")][0].split("[")[1].rstrip("]")
    scale_w = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit_w]
    summary = {"source": rep.split("/")[-1], "dram_bytes_per_launch": rd * scale + wr * scale_w, "kernels": result}
    json.dump(summary, open(out, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")

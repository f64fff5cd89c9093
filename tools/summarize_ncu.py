"""Text summaries of ncu exports for profiles/.

  summarize_ncu.py raw  <page-raw.csv>            key metrics per profiled kernel (from `ncu -i X.ncu-rep --page raw --csv`)
  summarize_ncu.py list <launch-list.csv>         per-kernel totals / shares of a `--metrics gpu__time_duration.sum` pass
"""
import collections
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__cycles_active.avg", "l1tex__cycles_elapsed.avg", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.max",
        "TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tc.sum", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"[{d['Kernel Name'][:100]}]  grid {d['Grid Size']} block {d['Block Size']}")
        for k in KEYS:
            if k in d and d[k] not in ("", "no data"):
                print(f"   {k:88s} {d[k]:>16s} {units[hdr.index(k)]}")
        stalls = [(float(d[k]), k) for k in hdr if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and d[k] not in ("", "no data")]
        for v, k in sorted(stalls, reverse=True)[:8]:
            name = k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")
            print(f"   stall cycles per issued instruction: {name:52s} {v:16.3f}")
        print()


def launch_list(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i
            break
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(list)
    for r in rows[start + 1:]:
        if len(r) > vi:
            name = r[ki].replace("void ", "").replace("rc::", "")
            agg[name[:72]].append(float(r[vi].replace(",", "")) / 1000.0)
    total = sum(sum(v) for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:72s} n={len(v):5d} total={sum(v):10.1f}us avg={sum(v) / len(v):8.2f}us share={100 * sum(v) / total:5.1f}%")


if __name__ == "__main__":
    {"raw": raw, "list": launch_list}[sys.argv[1]](sys.argv[2])

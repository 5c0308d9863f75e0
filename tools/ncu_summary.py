"""Key metrics + warp-stall breakdown of every kernel in an `ncu --page raw --csv` dump.
usage: ncu -i x.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv [title]"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size", "launch__cluster_dim_x",
        "launch__registers_per_thread", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum"]

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
if len(sys.argv) > 2:
    print("#", sys.argv[2])
for r in rows[2:]:
    print("kernel:", r[ix["Kernel Name"]][:110])
    for k in KEYS:
        if k in ix and r[ix[k]] != "":
            print(f"  {k:95s} {r[ix[k]]:>16s} {units[ix[k]]}")
    stalls = [(h.split("stalled_")[1].split("_per")[0], float(r[ix[h]])) for h in hdr
              if "issue_stalled" in h and "per_issue_active" in h and r[ix[h]] not in ("", "n/a")]
    tot = sum(v for _, v in stalls)
    print("  warp stalls per issued instruction (smsp__average_warps_issue_stalled_*_per_issue_active), total %.2f:" % tot)
    for n, v in sorted(stalls, key=lambda t: -t[1])[:8]:
        print(f"    {n:22s} {v:8.3f}  {100 * v / max(tot, 1e-9):5.1f} %")

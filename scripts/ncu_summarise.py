"""Dev tool: summarise an .ncu-rep (ncu --set full) into profiles/: a trimmed raw-page CSV of the libhlv
kernels and the traffic-vs-algorithmic-bytes entries bench.py reads for roofline.traffic.
  python scripts/ncu_summarise.py gpurun_out/prof_fused_r01.ncu-rep profiles/r01_ncu_full_fused.csv [--rows 13,38,63,88]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = 124_046_592
KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active")


def unit_scale(u):
    return {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(u, 1.0)


def main():
    rep, out_csv = sys.argv[1], sys.argv[2]
    rows_arg = None
    if "--rows" in sys.argv:
        rows_arg = [int(x) for x in sys.argv[sys.argv.index("--rows") + 1].split(",")]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rd = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rd[0], rd[1], rd[2:]
    cols = [i for i, h in enumerate(hdr) if h in ("ID", "Kernel Name") or h in KEEP]
    with open(out_csv, "w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow([hdr[i] for i in cols]); wr.writerow([units[i] for i in cols])
        for r in body:
            wr.writerow([r[i] for i in cols])
    ix = {h: i for i, h in enumerate(hdr)}
    summary = []
    for k, r in enumerate(body):
        def val(name):
            return float(r[ix[name]].replace(",", "")) * unit_scale(units[ix[name]])
        d = {"kernel": r[ix["Kernel Name"]][:90], "dram_bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
             "duration_ms_under_ncu": val("gpu__time_duration.sum"),
             "dram_pct_of_ncu_peak": float(r[ix["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]])}
        d["dram_gbs_under_ncu"] = d["dram_bytes"] / d["duration_ms_under_ncu"] / 1e6
        if rows_arg and k < len(rows_arg):
            d["rows"] = rows_arg[k]
        summary.append(d)
    print(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main()

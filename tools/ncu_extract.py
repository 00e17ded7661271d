"""Turn an `ncu --set full` report (gpurun_out/*.ncu-rep, scratch) into the small text summaries kept under
profiles/: selected raw metrics, the details page, the hottest source lines, and the JSON record bench.py reads
for `roofline` (warp instructions and DRAM bytes per launch of the dominant kernel).

  python tools/ncu_extract.py gpurun_out/prof.ncu-rep profiles/r2_pt_lookup_c4share --kernel amclj::pt_lookup_kernel \
         --particles 1250000 --beams 360 --workload "C4 per-GPU share"
"""
import argparse
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
METRICS = [
    "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__cycles_active.avg", "sm__cycles_active.min", "sm__cycles_active.max",
    "sm__cycles_elapsed.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]


def ncu(*args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out_prefix")
    ap.add_argument("--kernel", default=None)
    ap.add_argument("--particles", type=int, default=0)
    ap.add_argument("--beams", type=int, default=0)
    ap.add_argument("--workload", default="")
    a = ap.parse_args()
    raw = list(csv.reader(io.StringIO(ncu("-i", a.rep, "--page", "raw", "--csv"))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    row = dict(zip(hdr, zip(units, vals)))
    lines = [f"Kernel Name  {row.get('Kernel Name', ('', '?'))[1]}"]
    for m in METRICS:
        if m in row:
            lines.append(f"{m} {row[m][0]} {row[m][1]}")
    for h in hdr:
        if "issue_stalled" in h and h.endswith("per_warp_active.pct") and float(row[h][1] or 0) >= 2.0:
            lines.append(f"{h} {row[h][0]} {row[h][1]}")
    Path(a.out_prefix + "_raw_metrics.txt").write_text("\n".join(lines) + "\n")
    Path(a.out_prefix + "_details.txt").write_text(ncu("-i", a.rep, "--page", "details"))
    src = ncu("-i", a.rep, "--page", "source", "--csv", "--print-source", "cuda,sass")
    tmp = Path("/tmp/ncu_src.csv")
    tmp.write_text(src)
    out = subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_source_summary.py"), str(tmp), "lines", "40"],
                         capture_output=True, text=True).stdout
    Path(a.out_prefix + "_source_lines.txt").write_text(out)
    if a.kernel:
        def num(name, scale=1.0):
            u, v = row[name]
            v = float(v.replace(",", ""))
            mult = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9}.get(u, 1.0)
            return v * mult * scale
        rec = {"kernel": a.kernel, "workload": a.workload, "particles": a.particles, "beams": a.beams,
               "warp_instructions": num("smsp__inst_executed.sum"), "dram_bytes_read": num("dram__bytes_read.sum"),
               "dram_bytes_write": num("dram__bytes_write.sum"), "kernel_seconds_under_ncu": num("gpu__time_duration.sum"),
               "issue_active_pct": float(row["smsp__issue_active.avg.pct_of_peak_sustained_active"][1]),
               "alu_pipe_pct": float(row.get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", ("", "nan"))[1]),
               "fma_pipe_pct": float(row.get("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", ("", "nan"))[1]),
               "source": Path(a.out_prefix).name + "_raw_metrics.txt (ncu --set full --clock-control none, one launch)"}
        jf = ROOT / "profiles" / "dominant_kernel_metrics.json"
        data = json.loads(jf.read_text()) if jf.exists() else {"kernels": []}
        data["kernels"] = [k for k in data["kernels"] if not (k["kernel"] == rec["kernel"] and k["particles"] == rec["particles"]
                                                                and k["beams"] == rec["beams"])] + [rec]
        jf.write_text(json.dumps(data, indent=1) + "\n")
    print("\n".join(lines[:12]))


if __name__ == "__main__":
    main()

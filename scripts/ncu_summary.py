"""Condense gpurun_out/prof_<kernel>_<round>.ncu-rep (ncu --set full) into profiles/<kernel>_<round>.txt + profiles/traffic_<round>.json.
Run HERE (no GPU needed): python scripts/ncu_summary.py r02"""
import csv, io, json, os, subprocess, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_wait",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_lg_throttle",
        "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_branch_resolving",
        "smsp__pcsamp_warps_issue_stalled_membar", "smsp__pcsamp_warps_issue_stalled_mio_throttle", "smsp__pcsamp_warps_issue_stalled_sleeping"]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}
def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None
traffic = {}
for name in ("k_solve", "k_prep", "k_frame_cache_store", "k_knn_tc", "k_knn_select", "k_knn_rerank"):
    rep = os.path.join(ROOT, "gpurun_out", f"prof_{name}_{R}.ncu-rep")
    if not os.path.exists(rep):
        continue
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    lines = []
    for li, row in enumerate(rows[2:]):
        rec = dict(zip(hdr, row)); un = dict(zip(hdr, units))
        lines.append(f"kernel: {rec.get('Kernel Name')}   (launch {li} of this capture)")
        for m in WANT:
            if m in rec:
                lines.append(f"  {m:84s} {rec[m]} {un[m]}")
        if li == len(rows) - 3:
            b = sum(num(rec[m]) * UNIT.get(un[m], 1.0) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum") if m in rec)
            fp = [num(rec.get(f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum", "")) for op in ("ffma", "fmul", "fadd")]
            traffic[name] = {"dram_bytes_per_launch": b, "duration_us_under_ncu": num(rec["gpu__time_duration.sum"]) * (1e-3 if un["gpu__time_duration.sum"] == "ns" else 1.0),
                             "warp_inst_per_launch": num(rec.get("smsp__inst_executed.sum", "")),
                             "fp32_thread_inst_per_launch": (sum(v for v in fp if v is not None) if any(v is not None for v in fp) else None),
                             "fp32_flop_per_launch": (2 * fp[0] + fp[1] + fp[2]) if all(v is not None for v in fp) else None}
    open(os.path.join(ROOT, "profiles", f"{name}_{R}.txt"), "w").write("\n".join(lines) + "\n")
    print(name, traffic.get(name))
json.dump(traffic, open(os.path.join(ROOT, "profiles", f"traffic_{R}.json"), "w"), indent=1)

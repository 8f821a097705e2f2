"""Summarise an .ncu-rep (read here, no GPU needed) into the metrics DESIGN.md / bench.py quote.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [cells_per_launch] > profiles/<name>.txt"""
import csv, io, json, subprocess, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed_pipe_lsu.sum",
        "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_fmaheavy.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum"]


def to_bytes(val, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(val.replace(",", "")) * m.get(unit, 1)


def main():
    rep = sys.argv[1]
    cells = int(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# {rep}")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"\nkernel: {name}")
        vals = {}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                vals[w] = (r[i], units[i])
                print(f"  {w:75s} {r[i]:>18s} {units[i]}")
        if "dram__bytes_read.sum" in vals:
            rd = to_bytes(*vals["dram__bytes_read.sum"]); wr = to_bytes(*vals["dram__bytes_write.sum"])
            t = float(vals["gpu__time_duration.sum"][0].replace(",", ""))
            tu = {"us": 1e-6, "ns": 1e-9, "ms": 1e-3, "s": 1}[vals["gpu__time_duration.sum"][1]]
            print(f"  -> dram traffic per launch {rd + wr:.0f} B = {(rd + wr) / 1e6:.1f} MB ; {(rd + wr) / (t * tu) / 1e9:.0f} GB/s under ncu")
            if cells:
                print(f"  -> {(rd + wr) / cells:.1f} B per cell-update (read {rd / cells:.1f}, write {wr / cells:.1f})")


if __name__ == "__main__":
    main()

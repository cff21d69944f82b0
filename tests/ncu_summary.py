"""ncu_summary.py -- text summary of an .ncu-rep for profiles/ (run where ncu is installed, no GPU needed).
usage: python tests/ncu_summary.py report.ncu-rep [blocks per launch] > profiles/rNN_xxx_ncu.txt"""
import csv
import io
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rep = sys.argv[1]
    blocks = float(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print("# %s" % os.path.basename(rep))
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("\n== %s" % d.get("Kernel Name", "?"))
        for w in WANT:
            if w in d:
                print("   %-70s %s %s" % (w, d[w], u.get(w, "")))
        if blocks:
            try:
                t = float(d["gpu__time_duration.sum"])
                scale = {"ms": 1e6, "us": 1e3, "ns": 1.0, "s": 1e9}.get(u.get("gpu__time_duration.sum", "ms"), 1e6)
                print("   -> %.1f ns per stereo block, %.0f warp-instructions per block, %.0f DRAM bytes per block (%.0f blocks in the launch)"
                      % (t * scale / blocks, float(d["smsp__inst_executed.sum"]) / blocks,
                         (float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"])) *
                         {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(u.get("dram__bytes_read.sum", "Mbyte"), 1e6) / blocks, blocks))
            except Exception as e:
                print("   (per-block figures unavailable: %s)" % e)
        st = []
        for k in hdr:
            m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", k)
            if m and d.get(k) not in (None, "", "n/a"):
                try:
                    st.append((float(d[k]), m.group(1)))
                except ValueError:
                    pass
        if st:
            print("   warp stall reasons (warps per issue-active cycle):")
            for v, n in sorted(st, reverse=True)[:9]:
                print("      %-40s %.3f" % (n, v))


if __name__ == "__main__":
    main()

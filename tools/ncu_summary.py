"""Print the headline metrics of an .ncu-rep (read on the CPU box): python tools/ncu_summary.py file.ncu-rep"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active_realtime.avg.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread ", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct", "lts__t_bytes.sum ", "lts__t_sectors_op_write.sum ", "lts__t_sectors_op_read.sum ",
        "smsp__pcsamp_warps_issue_stalled", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct",
        "lts__t_sectors_srcunit_tex_op_write.sum ", "lts__t_sectors_srcunit_tex_op_read.sum ",
        "sm__inst_executed_pipe_uniform", "smsp__inst_executed.sum ", "dram__throughput"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("==", d.get("Kernel Name", "")[:100], d.get("Grid Size"), d.get("Block Size"))
        for h, u, v in zip(hdr, units, r):
            if any(k.strip() in h and (not k.endswith(" ") or h == k.strip()) for k in KEYS):
                if v not in ("0", "", "0.000000"):
                    print(f"   {h:90s} {v} {u}")


if __name__ == "__main__":
    main(sys.argv[1])

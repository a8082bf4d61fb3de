"""Key metrics of every launch in an ncu report (run where ncu is installed; no GPU needed):
python tools/ncu_key_metrics.py X.ncu-rep > profiles/X_ncu.txt"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    print(f"# {path}: ncu --set full --clock-control none (cold-cache replay, one launch per row)")
    for r in rows[2:]:
        print(f"\nkernel: {r[ki][:110]}")
        for h, u, v in zip(hdr, units, r):
            if h in WANT:
                print(f"  {h:86s} {v:>18s} {u}")


if __name__ == "__main__":
    main(sys.argv[1])

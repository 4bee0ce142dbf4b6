"""Extracts the metrics quoted in DESIGN.md / README.md from an `ncu --set full` report (read here, no GPU):

    python profiles/extract_full.py gpurun_out/r2_halo_gemm2_full.ncu-rep > profiles/r2_halo_gemm2_full.txt

The .ncu-rep itself stays in gpurun_out/ (scratch, not tracked)."""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"# {sys.argv[1]}: {len(data)} launches, ncu --set full --clock-control none --import-source on (raw page excerpt)")
    for name in WANT:
        if name not in hdr:
            continue
        i = hdr.index(name)
        print(f"{name} [{units[i]}]: " + " | ".join(r[i][:90] for r in data))


if __name__ == "__main__":
    main()

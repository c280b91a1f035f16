"""Print the roofline-relevant metrics of every kernel in an .ncu-rep (reads `ncu --page raw --csv`).
    python benchmarks/ncu_extract.py gpurun_out/x.ncu-rep [more.ncu-rep ...]"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__grid_size", "launch__cluster_dim_x",
        "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active"]

for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# {path}")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"## {d.get('Kernel Name')}")
        for k in WANT:
            if k in d:
                print(f"   {k:75s} {d[k]:>16s} {u.get(k, '')}")

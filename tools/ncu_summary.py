"""Key metrics of every kernel in an .ncu-rep (--set full) as text.  usage: python tools/ncu_summary.py report.ncu-rep"""
import csv, subprocess, sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "smsp__cycles_active.avg"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print(r[hdr.index("Kernel Name")][:160])
    for k in KEYS:
        if k in hdr:
            print(f"    {k:90s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    tot = sum(float(r[hdr.index(k)]) * scale.get(units[hdr.index(k)], 1.0)
              for k in ("dram__bytes_read.sum", "dram__bytes_write.sum") if k in hdr)
    print(f"    {'traffic = dram__bytes_read.sum + dram__bytes_write.sum':90s} {tot / 1e9:18.6f} GB")

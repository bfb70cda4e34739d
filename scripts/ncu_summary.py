"""Text summary of one .ncu-rep for profiles/: key raw metrics + hottest SASS.  usage: ncu_summary.py rep title"""
import csv, io, subprocess, sys
rep, title = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.max", "sm__inst_executed.sum.per_cycle_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]
print(f"# {title}\n# source: {rep} (B200, ncu --set full --clock-control none --import-source on)")
for i, h in enumerate(hdr):
    short = h.split(".", 2)[-1] if h.split(".")[0] in ("TPC", "SM_C", "LTS", "FBSP", "SM_A", "SM_B") else h
    if h in want or short in want:
        print(f"{short} = {vals[i]} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
open("/tmp/t/_src.csv", "w").write(src)
print("\n# hottest SASS by warp-stall samples")
print(subprocess.run([sys.executable, "scripts/ncu_hot.py", "/tmp/t/_src.csv", "14"], capture_output=True, text=True).stdout)

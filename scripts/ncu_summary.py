"""Summarise .ncu-rep captures (run where ncu is installed; no GPU needed):  python scripts/ncu_summary.py rep1 rep2 ..."""
import csv, subprocess, sys, io

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("sm__cycles_active.avg", "sm cycles"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu wavefronts %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("smsp__warps_active.avg.per_cycle_active", "warps/smsp"),
]

def summarise(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        print(f"== {path}\n   {d.get('Kernel Name','?')[:110]}")
        for k, label in KEYS:
            hit = [h for h in hdr if h == k or h.endswith("." + k)]
            if hit:
                print(f"   {label:20s} {d[hit[0]]} {u[hit[0]]}")

if __name__ == "__main__":
    for p in sys.argv[1:]:
        summarise(p)

"""Key raw metrics + stall breakdown of the first kernel in an ncu report.  usage: python tools/ncu_summary.py rep.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, unit, val = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, unit, val)}
print("kernel:", d.get("Kernel Name", ("?",))[0][:80])
for k in ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
          "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__warps_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
          "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
          "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
          "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
          "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
          "l1tex__data_bank_conflicts_pipe_lsu.sum"]:
    if k in d:
        print("%-75s %s %s" % (k, d[k][0], d[k][1]))
st = {k: float(v[0].replace(",", "")) for k, v in d.items()
      if k.startswith("smsp__pcsamp_warps_issue_stalled") and "not_issued" not in k}
tot = sum(st.values()) or 1
print("stalls: " + " | ".join("%s %.1f%%" % (k.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * v / tot)
                              for k, v in sorted(st.items(), key=lambda x: -x[1])[:12]))

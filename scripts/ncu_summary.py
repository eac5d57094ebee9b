"""Print the handful of ncu metrics we track from a .ncu-rep (raw page) + the top stall lines (source page)."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum.per_second',
        'l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.per_cycle_active', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__grid_size', 'launch__block_size', 'sm__cycles_active.avg', 'sm__cycles_elapsed.avg',
        'sm__cycles_elapsed.avg.per_second', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.avg.per_cycle_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum', 'lts__t_sectors_srcunit_tex_op_write.sum']
for h, u, v in zip(hdr, units, vals):
    if h in keep:
        print(f"{h}\t{u}\t{v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]; data = rows[2:]
iS, iSrc = hdr.index("# Samples"), hdr.index("Source")
cols = {k: hdr.index(k) for k in ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_barrier", "stall_math", "stall_mio", "stall_lg", "stall_not_selected", "stall_selected") if k in hdr}
data = [r for r in data if len(r) > iS and r[iS].isdigit()]
tot = sum(int(r[iS] or 0) for r in data)
print(f"# source page: {tot} warp-stall samples; totals by reason:", {k: sum(int(r[i] or 0) for r in data) for k, i in cols.items()})
for r in sorted(data, key=lambda r: -int(r[iS] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print(f"#  {r[iS]:>8}  {r[iSrc][:100]}")

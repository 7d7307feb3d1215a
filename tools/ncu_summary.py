"""Summarise an .ncu-rep: key raw metrics per kernel and the hottest source lines (cuda,sass view)."""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed_op_shared_atom.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'launch__shared_mem_per_block_dynamic', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for vals in rows[2:]:
    print("=" * 100)
    for h, u, v in zip(hdr, units, vals):
        if h in want:
            print("  %-70s %-10s %s" % (h, u, v[:120]))
    st = [(float(v), h) for h, v in zip(hdr, vals) if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio') and v not in ('', 'n/a')]
    print("  stalls (warps per issue):", ", ".join("%s %.2f" % (h.split('stalled_')[1].split('_per_issue')[0], x) for x, h in sorted(st, reverse=True)[:7]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur_file, cur_fn, data = None, None, {}
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == 'Function Name':
        cur_fn = r[1][:60]; continue
    if len(r) > 7 and r[0] not in ('', 'Line No') and r[2] == '-':
        try:
            data.setdefault(cur_fn, []).append((int(r[7]), int(r[6]), cur_file, int(r[0]), r[1][:105]))
        except ValueError:
            pass
for fn, d in data.items():
    tot = sum(x[0] for x in d) or 1; tots = sum(x[1] for x in d) or 1
    print("=" * 100); print(fn, "total warp-inst", tot, "samples", tots)
    for n, s, f, l, t in sorted(d, reverse=True)[:top]:
        print("%5.1f%% inst %5.1f%% smp %-14s:%-4d %s" % (100 * n / tot, 100 * s / tots, f, l, t))

"""Prints the handful of ncu metrics the design discussion uses from a .ncu-rep (first kernel in the report).
   python tools/ncu_summary.py report.ncu-rep"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(out)))
h, u, v = r[0], r[1], r[2]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'launch__shared_mem_per_block_dynamic', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_atom.sum']
for k in want:
    for i, n in enumerate(h):
        if n == k:
            print("%-70s %s %s" % (k, v[i], u[i]))
st = [(int(float(v[i])), n.replace('smsp__pcsamp_warps_issue_stalled_', '')) for i, n in enumerate(h)
      if n.startswith('smsp__pcsamp_warps_issue_stalled_') and not n.endswith('_not_issued')]
tot = sum(x for x, _ in st)
print("stall samples:", ", ".join("%s %.0f%%" % (n, 100.0 * x / tot) for x, n in sorted(st, reverse=True) if x * 50 > tot))

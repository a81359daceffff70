"""Summarises a .ncu-rep: for every captured launch the handful of metrics the design discussion uses.
   python tools/ncu_summary.py report.ncu-rep [--traffic workload]   (--traffic: merge dram bytes per launch into
   profiles/ncu_traffic.json under that workload name, keyed by the library's operation names)"""
import csv, subprocess, sys, io, json, os
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(out)))
h, u = r[0], r[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'launch__shared_mem_per_block_dynamic', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_atom.sum',
        'smsp__sass_inst_executed_op_global_red.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sectors_op_red.sum']
OPS = {"schur_mul_v4": "schur_multiply", "schur_mul_v3": "schur_multiply", "jtj_v4": "jtj_multiply", "jtj_v2": "jtj_multiply",
       "schur_init": "schur_init", "cam_blocks": "schur_diag_blocks", "evaluate_v": "evaluate_jacobian", "evaluate_kernel": "evaluate_jacobian",
       "backsub": "back_substitute", "model_cost": "model_cost", "cg_vector": "cg_vector", "row_q": "row_q"}
traffic = {}
ki = h.index('Kernel Name')
for v in r[2:]:
    name = v[ki]
    print("==== %s" % name[:110])
    vals = {}
    for k in want:
        if k in h:
            i = h.index(k)
            vals[k] = v[i]
            print("  %-68s %s %s" % (k, v[i], u[i]))
    st = [(float(v[i].replace(',', '') or 0), n.replace('smsp__pcsamp_warps_issue_stalled_', '')) for i, n in enumerate(h)
          if n.startswith('smsp__pcsamp_warps_issue_stalled_') and not n.endswith('_not_issued')]
    tot = sum(x for x, _ in st) or 1
    print("  stall samples:", ", ".join("%s %.0f%%" % (n, 100.0 * x / tot) for x, n in sorted(st, reverse=True) if x * 40 > tot))
    try:
        def num(key):
            i = h.index(key)
            x = float(v[i].replace(',', ''))
            unit = u[i].lower()
            return x * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
        b = num('dram__bytes_read.sum') + num('dram__bytes_write.sum')
        for pat, op in OPS.items():
            if pat in name:
                traffic[op] = b
    except Exception as e:
        pass
if "--traffic" in sys.argv:
    wl = sys.argv[sys.argv.index("--traffic") + 1]
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data.setdefault(wl, {}).update(traffic)
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)
    print("traffic ->", path, traffic)

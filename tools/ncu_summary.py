"""Per-kernel-kind summary of the raw page of the one-step ncu capture (tools/step_traffic.py):
  python tools/ncu_summary.py gpurun_out/r2_step_raw.csv > profiles/r2_step_ncu_full_summary.txt"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
names, units = rows[h], rows[h + 1]
col = {n: i for i, n in enumerate(names)}
want = ['gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
        'sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active', 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__inst_executed.sum', 'smsp__sass_inst_executed_op_local_ld.sum',
        'smsp__sass_inst_executed_op_local_st.sum', 'l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct']
stalls = [n for n in names if n.startswith('smsp__average_warps_issue_stalled') and n.endswith('per_issue_active.ratio')]
print("# ncu --set full --clock-control none, one un-graphed update of the bench population (256 agents, Ant-shaped), tools/step_traffic.py")
print("# source hash of csrc/ + include/saceo.h: see profiles/r2_step_traffic.json; selected metrics of the raw page, one launch per kernel kind")
print("# (utchmma ops path = tcgen05.mma throughput vs its peak; pipe_tensor_subpipe_hmma = mma.sync; local_ld / local_st = register spills;")
print("#  stall = warps stalled per issue, top 6)")
seen = set()
for r in rows[h + 2:]:
    if len(r) < len(names):
        continue
    kn = re.sub(r'\(.*$', '', r[col['Kernel Name']]).replace('saceo::', '').replace('void ', '')
    key = (kn, r[col['Grid Size']])
    if not kn.startswith(('k_mlp', 'k_model', 'k_adam', 'k_dw', 'k_gemm', 'k_gather')) or key in seen:
        continue
    seen.add(key)
    print("== %s grid %s block %s" % (kn, r[col['Grid Size']], r[col['Block Size']]))
    for w in want:
        if w in col and r[col[w]] not in ('', 'n/a'):
            print("  %-98s %s %s" % (w, r[col[w]], units[col[w]]))
    st = []
    for n in stalls:
        try:
            st.append((float(r[col[n]].replace(',', '')), n))
        except ValueError:
            pass
    for v, n in sorted(st, reverse=True)[:6]:
        print("  stall %6.2f %s" % (v, n.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))

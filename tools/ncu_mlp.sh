#!/bin/bash
# one `ncu --set full` capture of the fused policy forward (65,536 rows) -> gpurun_out/r2_mlp.ncu-rep + a few raw metrics
O=gpurun_out; mkdir -p $O
python tools/mlp_time.py 32768 > $O/mlp_plain.json 2>/dev/null || exit 1
ncu --set full --clock-control none --import-source on -k regex:evg_policy_mlp_kernel -s ${NCU_SKIP:-60} -c 1 -f -o $O/r2_mlp python tools/mlp_time.py 32768 > $O/r2_mlp_ncu.log 2>&1
ncu -i $O/r2_mlp.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); d=dict(zip(rows[0],rows[2]))
for k in ['gpu__time_duration.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor.sum','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','lts__t_bytes.sum','dram__bytes_read.sum','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','smsp__average_warp_latency_per_inst_issued.ratio']:
    print('%-75s %s'%(k,d.get(k)))
for k,v in sorted(d.items()):
    if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio') and float(v or 0)>0.3: print('   %-30s %s'%(k[34:-23],v))
for k,v in sorted(d.items()):
    if 'tensor' in k or 'tmem' in k.lower(): print('   ', k, v)
" | tee $O/r2_mlp_raw.txt

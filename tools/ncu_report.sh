#!/bin/bash
# tools/ncu_report.sh REPORT.ncu-rep LIB.so SOURCE.cu [kernel-substring]  ->  phase table, stall mix, hottest lines
set -e
rep=$(realpath $1); lib=$(realpath $2); src=$(realpath $3); k=${4:-evg_step_tpm_kernelILi11ELi12EhLi94ELb0ELi128ELb0E}
here=$(dirname $(realpath $0))
d=$(mktemp -d); cd $d
ncu -i $rep --page source --csv > sass.csv 2>/dev/null
ncu -i $rep --page raw --csv > raw.csv 2>/dev/null
cuobjdump -xelf all $lib > /dev/null
nvdisasm -g -c evg_step_tpm.sm_100a.cubin > k.sass 2>/dev/null
python - <<PY
import csv
rows=list(csv.reader(open('raw.csv')))
d=dict(zip(rows[0],rows[2]))
for k in ['gpu__time_duration.sum','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__icc_request_hit_rate.pct','dram__bytes_read.sum','dram__bytes_write.sum','smsp__average_warp_latency_per_inst_issued.ratio']:
    print('%-70s %s'%(k,d.get(k)))
for k,v in sorted(d.items()):
    if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio') and float(v or 0)>0.05:
        print('   %-30s %s'%(k[34:-23],v))
PY
python $here/ncu_phases.py sass.csv k.sass $k $src
python $here/ncu_lines.py sass.csv k.sass $k $src 2>&1 | sort -k5 -n -r | head -${TOPN:-25} | cut -c1-165
rm -rf $d

#!/bin/bash
# tools/make_profiles.sh TAG REPORT.ncu-rep LIB.so [LAUNCHES.csv]
# Distils an `ncu --set full` report of the step kernel (plus the launch list of a bench run) into the small text
# files kept under profiles/:  <TAG>_raw.csv (selected raw metrics), <TAG>_phases.txt, <TAG>_lines.txt,
# <TAG>_wavefronts.txt, <TAG>_launches.csv and step_kernel_traffic.json (read by bench.py for roofline.traffic).
set -e
tag=$1; rep=$(realpath $2); lib=$(realpath $3); launches=${4:+$(realpath $4)}
here=$(dirname $(realpath $0)); root=$(dirname $here); out=$root/profiles
src=$root/everglades-ai-wargame_b200/csrc/evg_step_tpm.cu
k=${KERNEL:-evg_step_tpm_kernelILi11ELi12EhLi94ELb0ELi128ELb0E}
matches=${MATCHES:-262144}
phase=${PHASE:-staggered}
d=$(mktemp -d); cd $d
ncu -i $rep --page source --csv > sass.csv 2>/dev/null
ncu -i $rep --page raw --csv > raw.csv 2>/dev/null
cuobjdump -xelf all $lib > /dev/null
nvdisasm -g -c evg_step_tpm.sm_100a.cubin > k.sass 2>/dev/null
python - "$out/${tag}_raw.csv" "$out/step_kernel_traffic.json" "$tag" "$matches" "$phase" <<'PY'
import csv, json, sys
rows = list(csv.reader(open('raw.csv')))
d = dict(zip(rows[0], zip(rows[1], rows[2])))
keep = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_st.sum',
        'sm__icc_request_hit_rate.pct', 'gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg',
        'smsp__average_warp_latency_per_inst_issued.ratio']
keep += sorted(k for k in d if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio'))
with open(sys.argv[1], 'w') as f:
    f.write('metric,unit,launch0\n')
    for k in keep:
        if k in d:
            f.write('%s,%s,%s\n' % (k, d[k][0], d[k][1]))
unit = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1.0}
rd = float(d['dram__bytes_read.sum'][1]) * unit[d['dram__bytes_read.sum'][0]]
wr = float(d['dram__bytes_write.sum'][1]) * unit[d['dram__bytes_write.sum'][0]]
m, phase = int(sys.argv[4]), sys.argv[5]
# one entry per (batch size, phase mix): bench.py reports roofline.traffic only from a capture of ITS batch size and phase
try:
    doc = json.load(open(sys.argv[2]))
    caps = [c for c in doc.get("captures", []) if not (int(c["envs"]) == m and c.get("phase") == phase)]
except Exception:
    caps = []
caps.append({"envs": m, "phase": phase, "kernel": d['Kernel Name'][1], "dram_bytes_read_per_launch": rd, "dram_bytes_write_per_launch": wr,
             "dram_bytes_per_launch": rd + wr, "dram_bytes_per_env_turn": (rd + wr) / m, "gpu_time_us": float(d['gpu__time_duration.sum'][1]),
             "source": "profiles/%s_raw.csv (ncu --set full --clock-control none, one launch of a settled `bench.py --phase %s --envs-per-gpu %d` run)" % (sys.argv[3], phase, m)})
json.dump({"what": "DRAM bytes of one step-kernel launch (dram__bytes_read.sum + dram__bytes_write.sum), per batch size and phase mix",
           "captures": sorted(caps, key=lambda c: (c["envs"], c["phase"]))}, open(sys.argv[2], 'w'), indent=1)
PY
python $here/ncu_phases.py sass.csv k.sass $k $src > $out/${tag}_phases.txt
python $here/ncu_lines.py sass.csv k.sass $k $src 2>&1 | head -80 | cut -c1-170 > $out/${tag}_lines.txt
TOPN=40 python $here/ncu_wavefronts.py sass.csv k.sass $k $src $matches > $out/${tag}_wavefronts.txt
if [ -n "$launches" ]; then grep -v "^==" $launches > $out/${tag}_launches.csv; fi
rm -rf $d
ls -la $out | grep $tag

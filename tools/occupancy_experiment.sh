#!/bin/bash
# step-kernel time of the thread-per-match kernel at 3, 2 and 1 resident CTAs per SM (extra dynamic smem)
for pad in 0 40000 120000; do
  EVG_TPM_SMEM_PAD=$pad python bench.py --steps 150 --warmup 150 --no-cpu-baseline --e2e-steps 2 2>/dev/null > /tmp/occ.json
  python -c "import json;d=json.load(open('/tmp/occ.json'));print('pad', $pad, 'kernel_ms', d['roofline']['kernel_ms'], 'ms_per_step', d['ms_per_step'])"
done

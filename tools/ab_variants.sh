#!/bin/bash
# On the GPU box: parity subset + the default bench for every build/libevgsim_*.so named on the command line
# (or all of them).  Output: gpurun_out/ab_<name>.json / .log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
libs=("$@"); [ ${#libs[@]} -eq 0 ] && libs=(build/libevgsim_*.so)
for lib in "${libs[@]}"; do
    name=$(basename "$lib" .so); name=${name#libevgsim_}
    export EVGSIM_LIB=$PWD/$lib
    if [ -z "$AB_SKIP_TESTS" ]; then
        timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_generic.py -x -q -m gpu > gpurun_out/ab_$name.log 2>&1
        echo "$name tests rc=$? $(tail -1 gpurun_out/ab_$name.log)"
    fi
    timeout 600 python bench.py --no-cpu-baseline --e2e-steps 1 ${AB_BENCH_ARGS} > gpurun_out/ab_$name.json 2>> gpurun_out/ab_$name.log
    python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/ab_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    print(sys.argv[1], "value %.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "kernel_ms", r.get("kernel_ms"), "frac", r.get("frac"))
except Exception as e:
    print(sys.argv[1], "bench failed:", e)
PY
done

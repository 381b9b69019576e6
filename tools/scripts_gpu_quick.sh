#!/bin/bash
# kernel + parity tests, then benches given as "name|args" lines in $BENCHES (newline separated)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -8; fi
echo "$BENCHES" | while IFS='|' read -r name args; do
  [ -z "$name" ] && continue
  MRD_BENCH_WATCHDOG=400 timeout 500 python bench.py --no-cpu-baseline $args --profile-out gpurun_out/prof_$name.csv > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err || { echo "$name FAILED"; tail -5 gpurun_out/bench_$name.err; continue; }
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]; d=json.load(open(f'gpurun_out/bench_{n}.json'))
print(n, 'value', round(d['value']), 'e2e', d['e2e'] and round(d['e2e']['value']), 'frac', round(d['tensor_peak_frac'],3), 'gemm TF', round(d['roofline']['achieved']), 'sm_mhz', d['clocks']['sm_mhz'], d['kernel_families'])
PY
done

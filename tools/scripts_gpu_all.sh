#!/bin/bash
# kernel tests + parity tests + a short bench with profile
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -15 | tee gpurun_out/tests.log
python bench.py --steps 3 --warmup 3 --global-batch ${GB:-1024} --no-cpu-baseline --profile-out gpurun_out/profile.csv ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "rc=$?"; tail -3 gpurun_out/bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value', round(d['value']), 'e2e', d['e2e'] and round(d['e2e']['value']), 'frac', round(d['tensor_peak_frac'],3), 'roofline', round(d['roofline']['achieved']), 'clocks', d['clocks'])
print(d['kernel_families'])
PY
cat gpurun_out/profile.csv | head -34

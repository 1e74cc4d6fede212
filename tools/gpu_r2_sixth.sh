#!/usr/bin/env bash
# round 2, call 6: packed quarter-list forward kernel, loss helper, new scan kernel: tests + A/B + bench
set -u
O=gpurun_out/r2f; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.txt
tail -n 12 $O/pytest_gpu.txt
GSR_TEST_OPTS=fwd_packed=0 timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "golden or live or oracle" > $O/pytest_fwd_scalar.txt 2>&1; tail -n 2 $O/pytest_fwd_scalar.txt
bash tools/gpu_ab_opts.sh "C3 full;C3 light;C2 light;C4 full;C4 light" "fwd_packed=1" "fwd_packed=0" 2>&1 | tee $O/ab.txt
timeout 900 python bench.py --no-extra > $O/bench_b200.json 2> $O/bench_b200.err; echo "bench exit $?"
python -c "
import json; a=json.load(open('$O/bench_b200.json'))
print(a['value'], a['ms_per_step'], 'e2e', a['e2e']['value'], a['e2e']['ms_per_step'], 'torch-loss e2e', a.get('e2e_torch_loss'), a.get('parity',{}).get('vs_reference'), a['stages_ms_per_step'], a['gpu_launches'])
"

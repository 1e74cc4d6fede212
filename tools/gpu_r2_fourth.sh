#!/usr/bin/env bash
# round 2, call 4: robust fast-exp quarter kernel, hoisted loads in the per-Gaussian kernels: tests + A/B + bench
set -u
O=gpurun_out/r2d; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.txt
tail -n 8 $O/pytest_gpu.txt
bash tools/gpu_ab_opts.sh "C3 full;C3 light;C2 light;C4 full;C1 light" "bwd_occ=0" "bwd_occ=7" "bwd_occ=8" 2>&1 | tee $O/ab.txt
timeout 900 python bench.py > $O/bench_b200.json 2> $O/bench_b200.err; echo "bench exit $?"
python -c "
import json; a=json.load(open('$O/bench_b200.json'))
print(a['value'], a['ms_per_step'], a['e2e']['value'], a.get('parity',{}).get('vs_reference'), a['stages_ms_per_step'])
for k,v in a.get('extra_configs',{}).items(): print(k, {q:v.get(q) for q in ('b200_fps','reference_fps','ratio','b200_e2e_fps','reference_e2e_fps','e2e_ratio')})
"

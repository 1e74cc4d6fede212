#!/usr/bin/env bash
# round 2, call 9 (2 GPUs): speculative forward blend (spec_render) + persistent SH rebuild under the per-Gaussian backward
set -u
O=gpurun_out/r2i; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.txt; tail -n 5 $O/pytest_gpu.txt
bash tools/gpu_ab_opts.sh "C2 light;C1 light;C3 full;C3 light" "spec_render=1" "spec_render=0" 2>&1 | tee $O/ab_spec.txt
python tools/host_profile.py --config C2 --variant light 2>&1 | head -n 3
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 \
     tools/exchange_bench.py --reps 12 --out $O/exchange_bench.jsonl > $O/xb_n2.log 2>&1; tail -n 1 $O/xb_n2.log | cut -c1-1800
for early in 1 0; do
  if [ $early = 0 ]; then export GSR_DP_NO_EARLY=1; else unset GSR_DP_NO_EARLY; fi
  for rep in a b; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$early \
     bench.py --gpus 2 --steps 20 --warmup 5 --cpu-frames 0 --no-extra --no-parity > $O/bench_n2_early$early$rep.json 2> $O/bench_n2_early$early$rep.err; echo "bench n2 early=$early exit $?"
  python - <<PY
import json
a=json.load(open("$O/bench_n2_early$early$rep.json"))
print("%.1f fps %.3f ms  e2e %.1f  exch %s  dp_check %s" % (a["value"], a["ms_per_step"], a["e2e"]["value"], a["stats"].get("exchange"), (a.get("dp_check") or {}).get("max_rel")))
PY
  done
done

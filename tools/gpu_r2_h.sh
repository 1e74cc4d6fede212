#!/usr/bin/env bash
# round 2, call 8 (2 GPUs): early masked-colour gather in the exchange (dp.py) + dp_check; host-side profile of the small frame
set -u
O=gpurun_out/r2h; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -q -x -k "arena or exchange or golden" > $O/pytest_subset.txt 2>&1; echo "pytest exit $?" >> $O/pytest_subset.txt; tail -n 3 $O/pytest_subset.txt
for early in 1 0; do
  if [ $early = 0 ]; then export GSR_DP_NO_EARLY=1; else unset GSR_DP_NO_EARLY; fi
  GSR_DP_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$early \
     bench.py --gpus 2 --steps 20 --warmup 5 --cpu-frames 0 --no-extra --no-parity > $O/bench_n2_early$early.json 2> $O/bench_n2_early$early.err; echo "bench n2 early=$early exit $?"
  python - <<PY
import json
a=json.load(open("$O/bench_n2_early$early.json"))
print("%.1f fps %.3f ms  e2e %.1f  exch %s  dp_check %s" % (a["value"], a["ms_per_step"], a["e2e"]["value"], a["stats"].get("exchange"), (a.get("dp_check") or {})))
PY
  grep -h "phases" $O/bench_n2_early$early.err
done
unset GSR_DP_NO_EARLY
python tools/host_profile.py --config C2 --variant light > $O/host_C2_light.txt 2>&1; head -n 60 $O/host_C2_light.txt
python tools/host_profile.py --config C2 --variant light --e2e > $O/host_C2_light_e2e.txt 2>&1; head -n 45 $O/host_C2_light_e2e.txt
python tools/host_profile.py --config C2 --variant light --impl reference > $O/host_C2_light_ref.txt 2>&1; head -n 4 $O/host_C2_light_ref.txt

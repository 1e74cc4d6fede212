#!/usr/bin/env bash
# round 2 (8 GPUs): the driver's scaling run with the final code: bench at 1, 2, 4 and 8 ranks (default exchange)
set -u
O=gpurun_out/r2s2; mkdir -p $O
for n in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2957$n \
     bench.py --gpus $n --steps 20 --warmup 5 --cpu-frames 0 --no-extra --no-parity > $O/bench_n$n.json 2> $O/bench_n$n.err; echo "bench n=$n exit $?"
  python - <<PY
import json
a=json.load(open("$O/bench_n$n.json"))
print("n=$n: %.1f fps %.3f ms  e2e %.1f (%.3f ms)  exch %s  dp_check %s" % (a["value"], a["ms_per_step"], a["e2e"]["value"], a["e2e"]["ms_per_step"], a["stats"].get("exchange"), (a.get("dp_check") or {}).get("max_rel")))
PY
done
python bench.py --steps 20 --warmup 5 --cpu-frames 0 --no-extra --no-parity > $O/bench_n1.json 2> $O/bench_n1.err
python -c "
import json; a=json.load(open('$O/bench_n1.json')); print('n=1:', a['value'], a['ms_per_step'], 'e2e', a['e2e']['value'])"

#!/usr/bin/env bash
# Scaling run on one 8-GPU box: N = 1, 2, 4, 8 back to back (view-DP, one all-reduce per step).
set -u
mkdir -p gpurun_out/scale
for n in 1 2 4 8; do
  if [ $n = 1 ]; then
    python bench.py --gpus 1 --steps 20 --warmup 5 --cpu-frames 0 > gpurun_out/scale/n$n.json 2> gpurun_out/scale/n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) \
      bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/scale/n$n.json 2> gpurun_out/scale/n$n.err
  fi
  python - <<PY
import json
a=json.load(open("gpurun_out/scale/n$n.json"))
print("N=$n: %.1f fps  %.3f ms/step  e2e %.1f fps" % (a["value"], a["ms_per_step"], a["e2e"]["value"]))
PY
done
grep -h "arena" gpurun_out/scale/*.err | head -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 \
  bench.py --gpus 8 --steps 20 --warmup 5 --dp-mode allreduce > gpurun_out/scale/n8_allreduce.json 2> gpurun_out/scale/n8_allreduce.err
python -c "
import json; a=json.load(open('gpurun_out/scale/n8_allreduce.json')); print('N=8 allreduce: %.1f fps  %.3f ms/step' % (a['value'], a['ms_per_step']))"

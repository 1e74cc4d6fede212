#!/usr/bin/env bash
set -u
O=gpurun_out/r2l; mkdir -p $O
python tools/e2e_phases.py --config C3 --variant full 2>&1 | tee $O/e2e_phases.txt
python tools/e2e_phases.py --config C2 --variant light --iters 300 2>&1 | tee -a $O/e2e_phases.txt
python tools/e2e_phases.py --config C3 --variant light 2>&1 | tee -a $O/e2e_phases.txt
python bench.py --steps 30 --warmup 5 --cpu-frames 0 --no-parity > $O/bench.json 2> $O/bench.err
python - <<PY
import json
a=json.load(open("$O/bench.json"))
print(a["value"], a["ms_per_step"], "e2e", a["e2e"]["value"], a["e2e"]["ms_per_step"], "torch-loss", a["e2e_torch_loss"]["value"])
for k,v in a.get('extra_configs',{}).items(): print(k, {q:(round(v.get(q),2) if isinstance(v.get(q),float) else v.get(q)) for q in ('b200_fps','reference_fps','ratio','b200_e2e_fps','reference_e2e_fps','e2e_ratio')})
PY

#!/usr/bin/env bash
# Bench lines for the other BASELINE.json configs (both arms), for the table in DESIGN.md / profiles.
set -u
mkdir -p gpurun_out/configs
for spec in "C2 light" "C2 full" "C1 light" "C4 full" "C4 light" "C3 light"; do
  set -- $spec
  python bench.py --config $1 --variant $2 --steps 20 --warmup 5 --cpu-frames 0 --no-extra --no-parity > gpurun_out/configs/b200_$1_$2.json 2>> gpurun_out/configs/err.log
  python bench.py --impl reference --config $1 --variant $2 --steps 5 --warmup 3 --no-extra --no-parity > gpurun_out/configs/ref_$1_$2.json 2>> gpurun_out/configs/err.log
  python - <<PY
import json
a=json.load(open("gpurun_out/configs/b200_$1_$2.json")); r=json.load(open("gpurun_out/configs/ref_$1_$2.json"))
print("$1 $2: ours %.1f fps (%.3f ms; e2e %.1f) ref %.1f fps (%.3f ms)  x%.1f   N=%s stages=%s" % (a["value"], a["ms_per_step"], a["e2e"]["value"], r["value"], r["ms_per_step"], a["value"]/r["value"], a["stats"]["num_rendered"], {k: round(v,3) for k,v in a.get("stages_ms_per_step",{}).items()}))
PY
done
tail -5 gpurun_out/configs/err.log

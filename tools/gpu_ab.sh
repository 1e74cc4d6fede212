#!/usr/bin/env bash
# A/B of a library option across configs: bash tools/gpu_ab.sh <option> "<configs>"
opt=$1; shift
for spec in "$@"; do
  set -- $spec
  for v in 0 1; do
    python bench.py --config $1 --variant $2 --steps 20 --warmup 5 --cpu-frames 0 --opt $opt=$v 2>/dev/null | python -c "
import json,sys; a=json.load(sys.stdin); s=a['stages_ms_per_step']
print('$1 $2 $opt=$v: %.1f fps  %.3f ms  N=%s  fwd %.3f bwd %.3f pre_b %.3f pre_f %.3f' % (a['value'], a['ms_per_step'], a['config']['num_rendered'], s['render_fwd'], s['render_bwd'], s['preprocess_bwd'], s['preprocess_fwd']))"
  done
done

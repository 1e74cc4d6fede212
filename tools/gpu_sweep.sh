#!/usr/bin/env bash
# Sweep one library option over a list of values: bash tools/gpu_sweep.sh <option> "<values>" "<config variant>"...
opt=$1; vals=$2; shift; shift
for spec in "$@"; do
  set -- $spec
  for v in $vals; do
    python bench.py --config $1 --variant $2 --steps 20 --warmup 5 --cpu-frames 0 --opt $opt=$v 2>/dev/null | python -c "
import json,sys; a=json.load(sys.stdin); s=a['stages_ms_per_step']
print('$1 $2 $opt=$v: %.1f fps  %.3f ms  fwd %.3f bwd %.3f pre_b %.3f pre_f %.3f scan %.3f emit %.3f sort %.3f' % (a['value'], a['ms_per_step'], s['render_fwd'], s['render_bwd'], s['preprocess_bwd'], s['preprocess_fwd'], s['scan'], s['emit_keys'], s['radix_sort']))"
  done
done

#!/usr/bin/env bash
# A/B of library option sets across configs:
#   bash tools/gpu_ab_opts.sh "<config variant>;<config variant>..." "<k=v,k=v>" "<k=v>" ...
IFS=';' read -ra SPECS <<< "$1"; shift
for spec in "${SPECS[@]}"; do
  set -- $spec "$@"; cfg=$1; var=$2; shift 2
  for optset in "$@"; do
    args=""; for kv in ${optset//,/ }; do args="$args --opt $kv"; done
    python bench.py --config $cfg --variant $var --steps 20 --warmup 5 --cpu-frames 0 --no-extra --no-parity $args 2>/dev/null | python -c "
import json,sys; a=json.load(sys.stdin); s=a['stages_ms_per_step']
print('$cfg $var [$optset]: %.1f fps  %.3f ms (p10 %.3f p90 %.3f)  e2e %.1f  fwd %.3f bwd %.3f pre_b %.3f pre_f %.3f scan %.3f scat %.3f sort %.3f memset %.3f launches/step %.1f' % (a['value'], a['ms_per_step'], a['step_ms']['p10_ms'], a['step_ms']['p90_ms'], a['e2e']['value'], s['render_fwd'], s['render_bwd'], s['preprocess_bwd'], s['preprocess_fwd'], s.get('tile_scan',0), s.get('tile_scatter',0), s.get('tile_sort',0), s.get('memset',0), a['gpu_launches']/a['steps']))"
  done
done

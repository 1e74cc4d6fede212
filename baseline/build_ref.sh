#!/usr/bin/env bash
# Build the UNMODIFIED reference (hjr37/diff-gaussian-rasterization, both variants) for sm_100
# into baseline/_ref/{light,full}. Sources are copied to a scratch dir (the reference tree is
# read-only), built there, and only the python package + the built _C .so are kept under
# baseline/_ref (git-ignored, travels to the GPU box with gpurun).
#
# The only source edit is the one SURVEY.md 8c records for the full variant: lines 203-347 of
# cuda_rasterizer/rasterizer_impl.cu are a stale copy of a legacy forward() that ends in a
# dangling "comment end*/" and does not compile; the live definition starts after it.
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
W="$(mktemp -d /tmp/refbuild.XXXXXX)"
export NVCC_APPEND_FLAGS="-include cstdint"
export TORCH_CUDA_ARCH_LIST="10.0"
export MAX_JOBS=${MAX_JOBS:-4}
build_one() {  # $1 = light|full
  local v=$1
  cp -r "$REF/diff-gaussian-rasterization-$v" "$W/$v"
  chmod -R u+w "$W/$v"
  rm -rf "$W/$v/build" "$W/$v"/*.egg-info "$W/$v"/diff_gaussian_rasterization/*.so "$W/$v"/diff_gaussian_rasterization/__pycache__
  if [ "$v" = full ]; then sed -i '203,347d' "$W/$v/cuda_rasterizer/rasterizer_impl.cu"; fi
  (cd "$W/$v" && python setup.py build_ext --inplace > "$W/$v.log" 2>&1) || { tail -50 "$W/$v.log"; exit 1; }
  mkdir -p "$OUT/$v"
  rm -rf "$OUT/$v/diff_gaussian_rasterization"
  cp -r "$W/$v/diff_gaussian_rasterization" "$OUT/$v/"
}
build_one light &
build_one full &
wait
ls -la "$OUT"/*/diff_gaussian_rasterization
rm -rf "$W"

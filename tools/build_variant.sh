#!/bin/bash
# Builds the library from the csrc/ tree of a git ref into build/libadsr_<name>.so, for A/B timing in ONE gpurun session:
#   tools/build_variant.sh HEAD base ;  ADSR_LIB=build/libadsr_base.so python tools/attn_block_bench.py
set -e
REF=${1:-HEAD}; NAME=${2:-base}; EXTRA=${3:-}     # REF = a git ref, or WORKTREE for the files as they are now; EXTRA = more nvcc flags
ROOT=$(cd "$(dirname "$0")/.." && pwd)
PKG=anomaly-detection-super-resolution_b200
D=$ROOT/build/variant_$NAME
rm -rf "$D"; mkdir -p "$D/$PKG" "$D/include"
if [ "$REF" = WORKTREE ]; then cp -r "$ROOT/$PKG/csrc" "$D/$PKG/"; cp "$ROOT"/include/*.h "$D/include/"; rm -f "$D/$PKG/csrc/"*.o
else git -C "$ROOT" archive "$REF" "$PKG/csrc" include | tar -x -C "$D"; fi
cd "$D/$PKG/csrc"
for f in *.cu; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $EXTRA -c -o "${f%.cu}.o" "$f" 2>/dev/null &
done
wait
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o "$ROOT/build/libadsr_$NAME.so" *.o
echo "built $ROOT/build/libadsr_$NAME.so from $REF"

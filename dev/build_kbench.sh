#!/bin/bash
# Builds dev/kbench (developer micro-benchmark): one object per kernel variant.
set -e
cd "$(dirname "$0")"
NV="nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -fmad=false"
mkdir -p _build
objs=""
i=0
while read -r tag flags; do
  [ -z "$tag" ] && continue
  case "$tag" in \#*) continue;; esac
  $NV -DVNAME=v$i -DVTAG="\"$tag\"" $flags -Xptxas -v -I. -c kvariant.cu -o _build/v$i.o 2> _build/v$i.log || { cat _build/v$i.log; exit 1; }
  echo "$tag: $(grep -A1 'rbis_fused_kernel' _build/v$i.log | grep -E 'spill' | head -1) $(grep -E 'Used [0-9]+ registers' _build/v$i.log | tail -1)"
  objs="$objs _build/v$i.o"
  i=$((i+1))
done < variants.txt
$NV kbench.cu $objs -o _build/kbench
echo built dev/_build/kbench

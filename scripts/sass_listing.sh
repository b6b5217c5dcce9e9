#!/bin/bash
# SASS evidence for profiles/: listing of the headline trace kernel (K=1 closest hit, SPREAD) and of the gather variant,
# with an opcode histogram and the mnemonics DESIGN.md refers to.  Runs in the GPU-less container (cuobjdump only).
set -e
cd "$(dirname "$0")/.."
OBJ=embree_b200/csrc/trace.o
OUT=profiles/${1:-r2}_trace_sass.txt
K='_ZN3rtk12trace_kernelILi1ELb0ELb0ELb0ELi0ELi0ELb1ELb0EEEvNS_11TraceParamsE'
G='_ZN3rtk12trace_kernelILi1ELb0ELb0ELb0ELi0ELi2ELb1ELb0EEEvNS_11TraceParamsE'
{
  echo "# cuobjdump -sass of embree_b200/csrc/trace.o (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3), $(date -u +%Y-%m-%d)"
  echo "# headline kernel: rtk::trace_kernel<K=1, OCCLUDED=false, STATS=false, ROBUST=false, GENERAL=0, GATHER=0, SPREAD=true, FILTER=false>"
  grep -A3 "$K" embree_b200/csrc/trace.o.ptxas.log | sed 's/^/# /'
  cuobjdump -sass -fun "$K" $OBJ | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4})\*\/\s+/\1  /; s/\s*\/\*.*$//' > /tmp/sass_k.txt
  echo "# instructions: $(wc -l < /tmp/sass_k.txt)"
  echo "# opcode histogram:"
  awk '{op=$2; if (op ~ /^@/) op=$3; sub(/\..*/, "", op); print op}' /tmp/sass_k.txt | sort | uniq -c | sort -rn | head -24 | sed 's/^/#   /'
  echo "# evidence mnemonics (count):"
  for m in "LDG.E.ENL2.256.CONSTANT" "UBLKPF" "STG.E.128" "LDS" "STS" "ATOMS" "REDUX" "SHFL" "VOTE" "I2F.U8" "LDL" "STL"; do
    echo "#   $m: $(grep -c "$m" /tmp/sass_k.txt || true)"
  done
  echo "# gather variant (GATHER=2): $(cuobjdump -sass -fun "$G" $OBJ | grep -cE 'STG.E.ENL2.256') x STG.E.ENL2.256 (256-bit stores of the staged hit records), $(cuobjdump -sass -fun "$G" $OBJ | grep -cE 'REDUX') x REDUX"
  echo
  cat /tmp/sass_k.txt
} > $OUT
echo "wrote $OUT ($(wc -l < $OUT) lines)"

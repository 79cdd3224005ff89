#!/bin/bash
# Variant libraries of ONE kernel source for A/B timing on the GPU box (tools/qp_time.py, tools/ppo_update_time.py ... with ML4CA_LIB=...):
#   tools/kernel_variants.sh qp_alloc.cu name1 "flags1" name2 "flags2" ...   ->  ml4ca_b200/libvar_<name>.so  (git-ignored, travels with gpurun)
# Only the named source is recompiled; the other objects come from the default build.
set -e
cd "$(dirname "$0")/.."
python -m ml4ca_b200.build > /dev/null
C=ml4ca_b200/csrc
SRC=$1; shift
BASE=${SRC%.cu}
OTHERS=$(ls $C/*.o | grep -v "/$BASE.o" | grep -v libvar)
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
       --expt-relaxed-constexpr -Xptxas -v $flags -c $C/$SRC -o /tmp/${BASE}_$name.o 2> /tmp/${BASE}_$name.log
  grep -E "spill" /tmp/${BASE}_$name.log | sort | uniq -c | sed "s/^/  [$name] /"
  nvcc -shared -o ml4ca_b200/libvar_$name.so /tmp/${BASE}_$name.o $OTHERS -lcudart
  echo built ml4ca_b200/libvar_$name.so "($flags)"
done

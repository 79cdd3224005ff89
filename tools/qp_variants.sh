#!/bin/bash
# Variant libraries of the QP allocator kernel for A/B timing on the GPU box (tools/qp_time.py with ML4CA_LIB=...):
#   tools/qp_variants.sh name1 "flags1" name2 "flags2" ...   ->  ml4ca_b200/libvar_<name>.so  (git-ignored, travels with gpurun)
# Only qp_alloc.cu is recompiled; the other objects come from the default build.
set -e
cd "$(dirname "$0")/.."
python -m ml4ca_b200.build > /dev/null
C=ml4ca_b200/csrc
OTHERS=$(ls $C/*.o | grep -v "qp_alloc" | grep -v libvar)
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
       --expt-relaxed-constexpr $flags -c $C/qp_alloc.cu -o /tmp/qp_alloc_$name.o
  nvcc -shared -o ml4ca_b200/libvar_$name.so /tmp/qp_alloc_$name.o $OTHERS -lcudart
  echo built ml4ca_b200/libvar_$name.so "($flags)"
done

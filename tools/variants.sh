#!/bin/bash
# Build libsesrq variants that differ in -D switches of sesrq_umma.cu only:  tools/variants.sh name1:"-DA -DB" name2:"-DC" ...
# Output: tools/_variants/libsesrq_<name>.so (select one at run time with SESRQ_LIB=<path>)
set -e
cd "$(dirname "$0")/../sesr_pytorch_quantize_b200/csrc"
OUT=../../tools/_variants; mkdir -p $OUT
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --fmad=false"
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  ( $NV $flags -c sesrq_umma.cu -o $OUT/umma_$name.o && \
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libsesrq_$name.so ../_build/sesrq_api.o ../_build/sesrq_generic.o $OUT/umma_$name.o ../_build/sesrq_chain.o ../_build/sesrq_calib.o ../_build/sesrq_ops.o ../_build/sesrq_fq.o -Xlinker --version-script=exports.map && rm $OUT/umma_$name.o && echo built $name ) &
done
wait

#!/bin/bash
# GPU box: time the default step with each variant library (tools/_variants/*.so) and the in-tree one.
# usage: tools/variant_bench.sh [bench args]   -> prints per-layer ms for outputs both / y / q5
run() {  # name libpath
  for o in ${OUTS:-both y q5}; do
    SESRQ_LIB=$2 timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --calib-frames 0 --no-variants --outputs $o "${@:3}" 2>>gpurun_out/variant_bench.err | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('%-16s %-5s step %.4f  layers %s'%('$1','$o',l['ms_per_step'],' '.join('%.4f'%x['avg_ms'] for x in l['layers'])))" || echo "$1 $o failed"
  done
}
run intree sesr_pytorch_quantize_b200/libsesrq.so "$@"
for f in tools/_variants/libsesrq_*.so; do n=$(basename $f .so); run ${n#libsesrq_} $f "$@"; done
run intree2 sesr_pytorch_quantize_b200/libsesrq.so "$@"

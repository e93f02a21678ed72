#!/bin/bash
# GPU box: exe_mode-0 parity tests + calibration-leg timing for the in-tree library and every tools/_variants/*.so
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-variants --no-workloads --no-python-ref"
for f in sesr_pytorch_quantize_b200/libsesrq.so tools/_variants/libsesrq_*.so; do
  echo "== $f"
  SESRQ_LIB=$f python -m pytest tests/test_fq_calib_gpu.py tests/test_myql_gpu.py -m gpu -x -q 2>&1 | tail -2
  for i in 1 2; do
  SESRQ_LIB=$f $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read())['calibration']; print('calib minmax %.2f ms  kl %.2f ms  hash %s bins %s' % (d['minmax_s']*1e3, d['kl_s']*1e3, d['minmax_threshold_hash'], d['kl_bin_counts']))"
  done
done

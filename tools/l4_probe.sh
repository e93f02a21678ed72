# final-layer store experiments: usage  OUT_LIST="both y q5" DBG_LIST="0 1 3 16 32" bash tools/l4_probe.sh
for o in ${OUT_LIST:-both y q5}; do for d in ${DBG_LIST:-0 1 2}; do
SESRQ_UMMA_DBG=$d timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --calib-frames 0 --outputs $o 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('outputs=$o dbg=$d L4 %.3f L3 %.3f'%(l['layers'][4]['avg_ms'], l['layers'][3]['avg_ms']))"
done; done

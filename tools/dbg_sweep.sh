# usage: DBG_LIST="0 3" NBW_LIST="1 2 3" TH_LIST="24 16" bash tools/dbg_sweep.sh
for th in ${TH_LIST:-99}; do for nbw in ${NBW_LIST:-0}; do for d in ${DBG_LIST:-0}; do
  [ "$nbw" != "0" ] && export SESRQ_UMMA_NBW=$nbw
  [ "$th" != "99" ] && export SESRQ_UMMA_MAXTH=$th
  SESRQ_UMMA_DBG=$d timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print('th=$th nbw=$nbw dbg=$d', 'us/frame %.1f'%l['stack']['us_per_frame'], ' '.join('L%d %.3f'%(x['layer'],x['avg_ms']) for x in l['layers']))"
done; done; done

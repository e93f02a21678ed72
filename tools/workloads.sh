# all BASELINE shapes on one GPU: usage  [WL="..."] bash tools/workloads.sh  (env knobs pass through)
for w in ${WL:-sesr_x2_int8_sim_1080p sesr_x4_int8_sim_540x960 sesr_x4_int8_sim_80x960 sesr_x2_int8_sim_80x960 nrdm3_int8_sim_80x960}; do
timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-e2e --calib-frames 0 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print('$w', '%.1f Gpix/s'%(l['value']/1e3), 'frac %.3f'%l['stack']['frac_of_hbm_peak'], 'us/frame %.2f'%l['stack']['us_per_frame'], ' '.join('L%d %.3f'%(x['layer'],x['avg_ms']) for x in l['layers']))"
done

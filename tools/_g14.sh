python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 3 --warmup 2 --only-extra config3_6x6mod_qroute,config4_lata_qroute,config4_lata_hybridq,config2_shortest_sweep 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for e in d['extra_configs']: print(e['key'], 'ms/step %.2f value %.4g e2e %.4g frac %.3f'%(e['ms_per_step'], e['value'], e['e2e']['value'], e['roofline']['frac']), e['launch'], e['step_ms'])"

# the four bench lines of the final build (profiles/r02_bench_final*.json)
python bench.py > gpurun_out/r02_final7.json 2> gpurun_out/r02_final7.err
for c in hr3d hr3d_one_hm hr3d_one_hm_doppler_phase; do python bench.py --cfg $c > gpurun_out/r02_final7_$c.json 2> gpurun_out/r02_final7_$c.err; done
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_final7*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d.get('ms_per_step'), d.get('value'), (d.get('e2e') or {}).get('value'), (d.get('inference') or {}).get('value'), d['roofline']['achieved'], d['roofline']['frac'])
    except Exception as e: print(f, 'ERR', e)
"

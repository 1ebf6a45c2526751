set -x
python bench.py > gpurun_out/r02_final6.json 2> gpurun_out/r02_final6.err
for c in hr3d hr3d_one_hm hr3d_one_hm_doppler_phase; do python bench.py --cfg $c > gpurun_out/r02_final6_$c.json 2> gpurun_out/r02_final6_$c.err; done
python bench.py --cfg hr3d_one_hm_doppler_phase --dcn-head --batch 4 > gpurun_out/r02_final6_dcn.json 2> gpurun_out/r02_final6_dcn.err
python bench.py --no-extras --timeline gpurun_out/timeline_final6.json > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_final6_ncu_launches_raw.csv python bench.py --steps 1 --warmup 3 --no-extras --no-graph --sync-wgrad --serial-branches > gpurun_out/r02_final6_ncu.log 2>&1
python -m pytest tests/test_full_grid_gpu.py -s -q > gpurun_out/r02_final6_parity.txt 2>&1
tail -3 gpurun_out/r02_final6_parity.txt
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_final6*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d.get('ms_per_step'), d.get('value'), (d.get('e2e') or {}).get('value'), (d.get('inference') or {}).get('value'))
    except Exception as e: print(f, 'ERR', e)
"

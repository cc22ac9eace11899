timeout -s KILL 60 python tools/mega_beacon.py 3 300 | tail -2
timeout -s KILL 120 python tools/mega_det.py 300 1 2 3 20 | cut -c1-150 | grep imax
timeout -s KILL 200 python tools/mega_check.py 100 | cut -c1-200
for k in 7 8 13; do JSTSP_DBG_KERNEL=$k timeout -s KILL 100 python tools/mega_probe.py | tail -6; done
timeout -s KILL 250 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_mega4.json 2> gpurun_out/bench_mega4.err; python -c "
import json; d=json.load(open('gpurun_out/bench_mega4.json')); print('value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'], 'nmse', d.get('nmse'), 'pipeline', d['pipeline']['value'])"

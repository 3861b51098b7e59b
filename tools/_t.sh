mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sql_gpu.py -m gpu -x -q -k "bins_head or decoder or tail" 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline --steps 200 2> gpurun_out/b1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'loss', d['loss'])
for k,v in d['kernels'].items():
    if 'head' in k: print('   %-28s %5.1f x %8.1f us'%(k, v['launches_per_step'], v['ms_per_step']/v['launches_per_step']*1e3))"
timeout 300 python bench.py --no-cpu-baseline --steps 200 --config 3 2> gpurun_out/b3.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c3', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'loss', d['loss'])
for k,v in d['kernels'].items():
    if 'head' in k: print('   %-28s %5.1f x %8.1f us'%(k, v['launches_per_step'], v['ms_per_step']/v['launches_per_step']*1e3))"

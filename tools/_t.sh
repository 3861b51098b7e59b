mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_photometric_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 200 python tools/time_photo.py 2>&1 | grep -E "photometric|ms_"
timeout 300 python bench.py --no-cpu-baseline --steps 200 2> gpurun_out/b1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'loss', d['loss'])"
timeout 300 python bench.py --no-cpu-baseline --steps 200 --config 3 2> gpurun_out/b3.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c3', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'loss', d['loss'])"

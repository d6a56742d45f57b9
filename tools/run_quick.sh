set -x
mkdir -p gpurun_out
T=${TAG:-r2f}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
SR_LOS_TIMING=1 python bench.py --pixels ${PIX:-3000} --steps 1 --warmup 1 --no-extras --no-cpu-baseline 2> gpurun_out/${T}_q.err | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'kern',json.dumps(d['kernels']),'roof',d['roofline']['frac'],'batch',json.dumps(d['batch']))"
grep plan: gpurun_out/${T}_q.err | tail -2

set -x
mkdir -p gpurun_out
T=${TAG:-r2h}
timeout 300 python -m pytest tests/test_gpu_los.py -m gpu -x -q -k fused_kernel > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_pytest.log
tail -15 gpurun_out/${T}_pytest.log
for v in 3 4; do
SR_LOS_VER=$v SR_LOS_TIMING=1 timeout 600 python bench.py --pixels ${PIX:-3000} --steps 1 --warmup 1 --no-extras --no-cpu-baseline 2> gpurun_out/${T}_q$v.err | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'kern',json.dumps(d['kernels']),'roof',d['roofline']['frac'],d['roofline']['kernel'],'batch',json.dumps(d['batch']))"
grep "plan" gpurun_out/${T}_q$v.err | tail -2
done

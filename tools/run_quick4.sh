mkdir -p gpurun_out
python -m pytest tests/test_gpu_los.py tests/test_gpu_jacobian.py tests/test_gpu_api.py -m gpu -q 2>&1 | tail -2
for t in 1 16; do
SR_LOS_PLAN_THREADS=$t SR_LOS_TIMING=1 python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline 2> gpurun_out/r2t_q.err | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],'chk',repr(d['batch']['checksum']))"
grep "plan:" gpurun_out/r2t_q.err | tail -1
done

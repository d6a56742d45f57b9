mkdir -p gpurun_out
python -m pytest tests/test_gpu_los.py tests/test_gpu_jacobian.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -2
for t in 1 2 4 8; do
echo "== TPC=$t"
SR_MMA_TPC=$t timeout 600 python bench.py --pixels 3000 --steps 1 --warmup 1 --no-extras --no-cpu-baseline 2> gpurun_out/r2r_q.err | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],'mma',d['kernels']['los_mma']['ms_per_step_per_gpu'],'layers',d['kernels']['los_layers']['ms_per_step_per_gpu'],'roof',d['roofline']['frac'],'chk',repr(d['batch']['checksum']))"
done

mkdir -p gpurun_out
python -m pytest tests/test_gpu_los.py tests/test_gpu_slab.py tests/test_gpu_api.py tests/test_gpu_sza.py -m gpu -q 2>&1 | tail -3
python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline 2> gpurun_out/r2m_q.err | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],'kern',json.dumps(d['kernels']),'roof',d['roofline']['frac'],'chk',repr(d['batch']['checksum']))"

mkdir -p gpurun_out
python -m pytest tests/test_gpu_api.py tests/test_gpu_sza.py tests/test_gpu_examples.py tests/test_gpu_configs.py -m gpu -q 2>&1 | tail -3
python bench.py --steps 2 --warmup 2 --no-extras --no-cpu-baseline 2> gpurun_out/r2p_q.err | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],'kern',json.dumps(d['kernels']),'roof',d['roofline']['frac'],'chk',repr(d['batch']['checksum']),'steps_s',d['batch']['steps_build_and_gather_s'])"

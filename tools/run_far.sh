mkdir -p gpurun_out
python -m pytest tests/test_gpu_voigt.py tests/test_gpu_ref_golden.py tests/test_gpu_slab.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -12
for f in 0 1; do echo "== FAR=$f"; SR_K1_FAR=$f python tools/prof_run.py k1b 16 2>&1 | tail -1; done

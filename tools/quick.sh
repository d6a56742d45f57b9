mkdir -p gpurun_out
python -m pytest tests/test_gpu_los.py -x -q 2>&1 | tail -2
SR_PROF_NLOS=36 python tools/prof_run.py k3
SR_PROF_NLOS=96 python tools/prof_run.py fused

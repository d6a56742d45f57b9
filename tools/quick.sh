mkdir -p gpurun_out
python -m pytest tests/test_gpu_los.py tests/test_gpu_api.py -x -q 2>&1 | tail -2
echo OVERLAP; SR_PROF_NLOS=96 python tools/prof_run.py fused
echo OVERLAP_MINB3; SR_MMA_MINB=3 SR_PROF_NLOS=96 python tools/prof_run.py fused
echo NOOVERLAP; SR_LOS_NOOVERLAP=1 SR_PROF_NLOS=96 python tools/prof_run.py fused
echo OVERLAP_360; SR_PROF_NLOS=360 python tools/prof_run.py fused

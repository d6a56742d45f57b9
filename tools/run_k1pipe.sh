set -x
mkdir -p gpurun_out
T=${TAG:-r2b}
L=gpurun_out/${T}_k1pipe.log
: > $L
python -m pytest tests/test_gpu_voigt.py tests/test_gpu_api.py -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
for nc in 8 16 32; do
  echo "== PIPE=0 n_c=$nc" >> $L; SR_K2_PIPE=0 python tools/prof_run.py k1b $nc >> $L 2>&1
  for sub in 2 4 8; do
    echo "== PIPE=1 SUB=$sub n_c=$nc" >> $L; SR_K2_PIPE=1 SR_K2_SUB=$sub python tools/prof_run.py k1b $nc >> $L 2>&1
  done
done
grep -E "==|k1b" $L

set -x
mkdir -p gpurun_out
T=${TAG:-r2a}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
export SR_PROF_NPIX=1000
L=gpurun_out/${T}_l2exp.log
: > $L
echo "== default" >> $L; python tools/prof_run.py batch >> $L 2>&1
for cfg in "512 192" "1024 96" "2048 48" "1024 192" "4096 48" "512 384"; do
  set -- $cfg
  for keep in 0 1; do
    echo "== CHUNK=$1 BLOCK=$2 KEEP=$keep" >> $L
    SR_LOS_CHUNK=$1 SR_LOS_BLOCK=$2 SR_LOS_L2KEEP=$keep python tools/prof_run.py batch >> $L 2>&1
  done
done
echo "== CHUNK=1024 BLOCK=96 KEEP=1 LD=1" >> $L
SR_MMA_LD=1 SR_LOS_CHUNK=1024 SR_LOS_BLOCK=96 SR_LOS_L2KEEP=1 python tools/prof_run.py batch >> $L 2>&1
grep -E "==|batch" $L

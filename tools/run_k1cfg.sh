set -x
mkdir -p gpurun_out
T=${TAG:-r2c}
L=gpurun_out/${T}_k1cfg.log
: > $L
for cfg in 5 11 12 13 14 15 0; do
  echo "== CFG=$cfg" >> $L; SR_K1_CFG=$cfg python tools/prof_run.py k1b 8 >> $L 2>&1
done
grep -E "==|k1b" $L

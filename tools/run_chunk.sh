set -x
mkdir -p gpurun_out
T=${TAG:-r2e}
python -m pytest tests/test_gpu_slab.py -m gpu -x -q 2>&1 | tail -3
L=gpurun_out/${T}_chunk.log
: > $L
for ch in 65536 65280 61440 49152 98304 98560 32768 33024; do
  echo "== CHUNK=$ch" >> $L
  SR_LOS_TIMING=1 SR_LOS_CHUNK=$ch python bench.py --pixels 3000 --steps 1 --warmup 1 --no-extras --no-cpu-baseline 2>> $L | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],'kern',json.dumps(d['kernels']),'roof',d['roofline']['frac'])" >> $L
done
grep -E "==|value|plan: 9000" $L | uniq

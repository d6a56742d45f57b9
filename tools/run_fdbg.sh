mkdir -p gpurun_out
T=${TAG:-r2i}
for dbg in ${DBGS:-0 1 2 3}; do
echo "== FDBG=$dbg $EXTRA"
SR_LOS_FDBG=$dbg SR_LOS_VER=4 SR_LOS_TIMING=1 timeout 600 python bench.py --pixels ${PIX:-3000} --steps 1 --warmup 1 --no-extras --no-cpu-baseline 2> gpurun_out/${T}_q.err | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],'kern',json.dumps(d['kernels']),'roof',d['roofline']['frac'])"
grep "plan" gpurun_out/${T}_q.err | tail -1
done

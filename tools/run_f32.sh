mkdir -p gpurun_out
T=${TAG:-r2j}
for cfg in "0 4" "1 2" "1 4" "1 8"; do
set -- $cfg
echo "== F32=$1 UNROLL=$2"
SR_K3F_UNROLL=$2 SR_LOS_F32=$1 SR_LOS_VER=3 SR_LOS_TIMING=1 timeout 600 python bench.py --pixels ${PIX:-3000} --steps 1 --warmup 1 --no-extras --no-cpu-baseline 2> gpurun_out/${T}_q.err | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],'kern',json.dumps(d['kernels']),'roof',d['roofline']['frac'],'chk',repr(d['batch']['checksum']))"
grep "plan" gpurun_out/${T}_q.err | tail -1
done

mkdir -p gpurun_out
for mb in 8192 16384 24576; do
echo "== RADCAP_MB=$mb"
SR_LOS_RADCAP_MB=$mb SR_LOS_TIMING=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline 2> gpurun_out/r2l_q.err | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],'kern',json.dumps(d['kernels']),'roof',d['roofline']['frac'],'chk',repr(d['batch']['checksum']))"
grep "plan" gpurun_out/r2l_q.err | tail -1
done

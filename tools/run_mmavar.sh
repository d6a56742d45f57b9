mkdir -p gpurun_out
for cfg in "4 0" "4 1" "4 2" "3 0" "3 2"; do
set -- $cfg
echo "== MINB=$1 LD=$2"
SR_MMA_MINB=$1 SR_MMA_LD=$2 timeout 600 python bench.py --pixels 3000 --steps 1 --warmup 1 --no-extras --no-cpu-baseline 2> gpurun_out/r2o_q.err | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],'mma',d['kernels']['los_mma']['ms_per_step_per_gpu'],'layers',d['kernels']['los_layers']['ms_per_step_per_gpu'],'roof',d['roofline']['frac'])"
done

mkdir -p gpurun_out
T=r2q
echo skip pytest
echo skip smoke
SECONDS=0; python3 bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err; echo "ref rc $?"; echo "ref wall $SECONDS s"
SECONDS=0; python3 bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc $?"; echo "bench wall $SECONDS s"
python - <<PY
import json
r=json.loads(open("gpurun_out/${T}_ref.json").read().strip().splitlines()[-1])
d=json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
print("ref value",r["value"],r["cpu_baseline"]["cores"],"same config",r["config"]==d["config"])
print("ours value",d["value"],"e2e",d["e2e"]["value"],"ms",d["ms_per_step"],"ratio",d["value"]/r["value"],"e2e ratio",d["e2e"]["value"]/r["value"])
print("roofline",json.dumps(d["roofline"])[:900])
print("clocks",d["clocks"],"launches",d["gpu_launches"])
print("len json",len(json.dumps(d)))
print(json.dumps(d)[-1500:])
PY

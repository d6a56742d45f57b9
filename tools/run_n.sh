mkdir -p gpurun_out
T=${TAG:-r2k}
N=${N:-8}
if [ "$N" = "1" ]; then
python bench.py --gpus 1 --steps ${STEPS:-2} --warmup 3 > gpurun_out/${T}_n1.json 2> gpurun_out/${T}_n1.err; echo "n1 rc $?"
else
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-2} --warmup 3 > gpurun_out/${T}_n$N.json 2> gpurun_out/${T}_n$N.err; echo "n$N rc $?"
fi
tail -3 gpurun_out/${T}_n$N.err
python - <<PY
import json
f="gpurun_out/${T}_n$N.json"
d=json.loads(open(f).read().strip().splitlines()[-1])
print(f,"value",d["value"],"e2e",d["e2e"]["value"],"ms",d["ms_per_step"],"chk",repr(d["batch"]["checksum"]),"steps_s",d["batch"]["steps_build_and_gather_s"])
print("  lut_build",d["lut_build"]["value"],d["lut_build"]["roofline_frac_fp64"],"voigt",d["voigt"]["value"],d["voigt"]["roofline"]["frac"],"1e6",d["voigt_1e6"]["value"],d["voigt_1e6"]["ms"],"k3",d["k3_layers"]["value"],d["k3_layers"]["roofline"]["frac"],"jac",d["jacobian"]["value"],"hires",d["hires_host"]["value"])
print("  kern",json.dumps(d["kernels"]),"roof",d["roofline"]["frac"],d["roofline"]["share_of_step"])
print("  cpu",d.get("cpu_baseline",{}).get("value"), "clocks", d["clocks"])
PY

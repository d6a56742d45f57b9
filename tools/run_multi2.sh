set -x
mkdir -p gpurun_out
T=${TAG:-r2g}
N=${N:-2}
python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_los.py tests/test_gpu_jacobian.py -m gpu -x -q 2>&1 | tail -3
python bench.py --pixels ${PIX:-3000} --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_n1.json 2> gpurun_out/${T}_n1.err; echo "n1 rc $?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --pixels ${PIX:-3000} --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_n$N.json 2> gpurun_out/${T}_n$N.err; echo "n$N rc $?"
tail -5 gpurun_out/${T}_n$N.err
python - <<PY
import json
for f in ("gpurun_out/${T}_n1.json","gpurun_out/${T}_n$N.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,"unreadable",e); continue
    print(f,"value",d["value"],"e2e",d["e2e"]["value"],"ms",d["ms_per_step"],"chk",repr(d["batch"]["checksum"]),"steps_s",d["batch"]["steps_build_and_gather_s"])
    print("  lut_build",d["lut_build"]["value"],"voigt",d["voigt"]["value"],d["voigt"]["roofline"]["frac"],"1e6",d["voigt_1e6"]["value"],d["voigt_1e6"]["ms"],"k3",d["k3_layers"]["value"],d["k3_layers"]["roofline"]["frac"],"jac",d["jacobian"]["value"],"hires",d["hires_host"]["value"])
    print("  kern",json.dumps(d["kernels"]),"roof",d["roofline"]["frac"],d["roofline"]["share_of_step"])
PY

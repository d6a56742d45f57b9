set -x
mkdir -p gpurun_out
T=${TAG:-r2d}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
python bench.py --small --steps 2 --warmup 1 > gpurun_out/${T}_small.json 2> gpurun_out/${T}_small.err; echo "small rc $?"
tail -3 gpurun_out/${T}_small.err
SR_LOS_TIMING=1 python bench.py --steps ${STEPS:-2} --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc $?"
tail -5 gpurun_out/${T}_bench.err
cat gpurun_out/${T}_bench.json | tail -c 3000

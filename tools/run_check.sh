set -x
mkdir -p gpurun_out
T=${TAG:-s12}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc $?"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv

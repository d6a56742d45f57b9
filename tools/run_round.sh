set -x
mkdir -p gpurun_out
T=${TAG:-s5}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc $?"
SR_PROF_NLOS=96 python tools/prof_run.py fused > gpurun_out/${T}_fused_plain.log 2>&1
SR_PROF_NLOS=96 ncu --set full --clock-control none --import-source on -k regex:k_los_mma -s 2 -c 1 -o gpurun_out/${T}_mma -f python tools/prof_run.py fused > gpurun_out/${T}_mma_ncu.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv

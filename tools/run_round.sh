set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/s2_pytest.log
python bench.py > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err; echo "bench rc $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s2_bench_ref.json 2>> gpurun_out/s2_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s2_ncu_bench.log 2>&1
python tools/prof_run.py k1 > gpurun_out/s2_k1_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_voigt_tile -s 4 -c 1 -o gpurun_out/s2_k1 -f python tools/prof_run.py k1 > gpurun_out/s2_k1_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_los_layers -s 2 -c 1 -o gpurun_out/s2_k3 -f python tools/prof_run.py k3 > gpurun_out/s2_k3_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_los_fused -s 1 -c 1 -o gpurun_out/s2_fused -f python tools/prof_run.py fused > gpurun_out/s2_fused_ncu.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc

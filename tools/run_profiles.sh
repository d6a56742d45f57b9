# final ncu evidence pass of a round: every command first runs plain (exit 0), then under ncu
set -x
mkdir -p gpurun_out
T=${TAG:-r1}
python bench.py --steps 2 --warmup 3 --batch-pixels 120 --no-cpu-baseline > gpurun_out/${T}_bench_prof_plain.json 2> gpurun_out/${T}_bench_prof_plain.err; echo "plain rc $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${T}_bench_launches.csv python bench.py --steps 2 --warmup 3 --batch-pixels 120 --no-cpu-baseline > gpurun_out/${T}_bench_ncu.json 2> gpurun_out/${T}_bench_ncu.err; echo "ncu launches rc $?"
SR_PROF_NLOS=36 python tools/prof_run.py k3 > gpurun_out/${T}_k3_plain.log 2>&1
SR_PROF_NLOS=36 ncu --set full --clock-control none --import-source on -k regex:k_los_layers -s 2 -c 1 -o gpurun_out/${T}_k3 -f python tools/prof_run.py k3 > gpurun_out/${T}_k3_ncu.log 2>&1
SR_PROF_NPIX=100 python tools/prof_run.py batchjac > gpurun_out/${T}_batchjac_plain.log 2>&1
SR_PROF_NPIX=100 ncu --set full --clock-control none --import-source on -k regex:k_los_layers_jac -s 2 -c 1 -o gpurun_out/${T}_jac -f python tools/prof_run.py batchjac > gpurun_out/${T}_jac_ncu.log 2>&1
SR_PROF_NPIX=100 ncu --set full --clock-control none --import-source on -k regex:k_convolve_lowres -s 2 -c 1 -o gpurun_out/${T}_conv -f python tools/prof_run.py batchjac > gpurun_out/${T}_conv_ncu.log 2>&1
SR_PROF_NPIX=100 ncu --set full --clock-control none --import-source on -k regex:k_los_mma -s 2 -c 1 -o gpurun_out/${T}_mma -f python tools/prof_run.py batchjac > gpurun_out/${T}_mma_ncu.log 2>&1
SR_PROF_NPIX=1000 ncu --set full --clock-control none --import-source on -k regex:k_steps_ -s 2 -c 2 -o gpurun_out/${T}_steps -f python tools/prof_run.py batch > gpurun_out/${T}_steps_ncu.log 2>&1
python tools/prof_run.py k1b > gpurun_out/${T}_k1b_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_voigt_tile -s 3 -c 1 -o gpurun_out/${T}_tile -f python tools/prof_run.py k1b > gpurun_out/${T}_tile_ncu.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv

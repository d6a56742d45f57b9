# round-2 ncu evidence pass: every command first runs plain (exit 0), then under ncu
set -x
mkdir -p gpurun_out
T=r2
B="python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline"
# (1) launch list of the bench command with a shortened batch
$B --pixels 600 > gpurun_out/${T}_bench_prof_plain.json 2> gpurun_out/${T}_bench_prof_plain.err; echo "plain rc $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${T}_bench_launches.csv $B --pixels 600 > gpurun_out/${T}_bench_ncu.json 2> gpurun_out/${T}_bench_ncu.err; echo "ncu launches rc $?"
# (2) the dominant kernel on the bench's own launch shape (full batch: LOS blocks of 894 x 65536 points)
python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${T}_full_plain.json 2> gpurun_out/${T}_full_plain.err; echo "full plain rc $?"
ncu --set full --clock-control none --import-source on -k regex:k_los_mma -s 700 -c 1 -o gpurun_out/${T}_mma -f python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${T}_mma_ncu.log 2>&1; echo "mma rc $?"
ncu --set full --clock-control none --import-source on -k regex:k_los_layers_f32 -s 700 -c 1 -o gpurun_out/${T}_layers -f python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${T}_layers_ncu.log 2>&1; echo "layers rc $?"
# (3) the fused kernel (selectable), 3000 pixels
SR_LOS_VER=4 $B --pixels 3000 > gpurun_out/${T}_fused_plain.json 2> gpurun_out/${T}_fused_plain.err; echo "fused plain rc $?"
SR_LOS_VER=4 ncu --set full --clock-control none --import-source on -k regex:k_los_fused2 -s 14 -c 1 -o gpurun_out/${T}_fused -f $B --pixels 3000 > gpurun_out/${T}_fused_ncu.log 2>&1; echo "fused rc $?"
# (4) K1
python tools/prof_run.py k1b 16 > gpurun_out/${T}_k1b_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_voigt_tile -s 3 -c 1 -o gpurun_out/${T}_tile -f python tools/prof_run.py k1b 16 > gpurun_out/${T}_tile_ncu.log 2>&1; echo "tile rc $?"
ncu --set full --clock-control none --import-source on -k regex:k_core_eval -s 3 -c 1 -o gpurun_out/${T}_core -f python tools/prof_run.py k1b 16 > gpurun_out/${T}_core_ncu.log 2>&1; echo "core rc $?"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
ls -la gpurun_out/${T}_*.ncu-rep

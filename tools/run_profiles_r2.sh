# round-2 ncu evidence pass: every command first runs plain (exit 0), then under ncu; the .ncu-rep
# files are reduced to text on the box (gpurun_out/ is limited to 64 MiB)
set -x
mkdir -p gpurun_out
T=r2
B="python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline"
summ() {   # $1 = report stem
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/$1_raw.csv > gpurun_out/$1_summary.txt 2>&1
  ncu -i gpurun_out/$1.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | gzip > gpurun_out/$1_source.csv.gz
  rm -f gpurun_out/$1.ncu-rep gpurun_out/$1_raw.csv
}
if [ -z "$SKIP_LIST" ]; then
$B --pixels 600 > gpurun_out/${T}_bench_prof_plain.json 2> gpurun_out/${T}_bench_prof_plain.err; echo "plain rc $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${T}_bench_launches.csv $B --pixels 600 > gpurun_out/${T}_bench_ncu.json 2> gpurun_out/${T}_bench_ncu.err; echo "ncu launches rc $?"
fi
# the dominant kernels on the bench's own launch shape (LOS blocks of 894 x 65536 points)
$B --pixels 3000 > gpurun_out/${T}_b3000_plain.json 2> gpurun_out/${T}_b3000_plain.err; echo "b3000 plain rc $?"
ncu --set full --clock-control none --import-source on -k regex:k_los_mma -s 60 -c 1 -o gpurun_out/${T}_mma -f $B --pixels 3000 > gpurun_out/${T}_mma_ncu.log 2>&1; echo "mma rc $?"; summ ${T}_mma
ncu --set full --clock-control none --import-source on -k regex:k_los_layers_f32 -s 60 -c 1 -o gpurun_out/${T}_layers -f $B --pixels 3000 > gpurun_out/${T}_layers_ncu.log 2>&1; echo "layers rc $?"; summ ${T}_layers
# the fused kernel (selectable), 3000 pixels
SR_LOS_VER=4 $B --pixels 3000 > gpurun_out/${T}_fused_plain.json 2> gpurun_out/${T}_fused_plain.err; echo "fused plain rc $?"
SR_LOS_VER=4 ncu --set full --clock-control none --import-source on -k regex:k_los_fused2 -s 14 -c 1 -o gpurun_out/${T}_fused -f $B --pixels 3000 > gpurun_out/${T}_fused_ncu.log 2>&1; echo "fused rc $?"; summ ${T}_fused
# K1
python tools/prof_run.py k1b 16 > gpurun_out/${T}_k1b_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_voigt_tile -s 3 -c 1 -o gpurun_out/${T}_tile -f python tools/prof_run.py k1b 16 > gpurun_out/${T}_tile_ncu.log 2>&1; echo "tile rc $?"; summ ${T}_tile
ncu --set full --clock-control none --import-source on -k regex:k_core_eval -s 3 -c 1 -o gpurun_out/${T}_core -f python tools/prof_run.py k1b 16 > gpurun_out/${T}_core_ncu.log 2>&1; echo "core rc $?"; summ ${T}_core
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
du -sh gpurun_out; ls -la gpurun_out

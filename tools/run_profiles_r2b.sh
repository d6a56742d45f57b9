# K1 after the far-field path: plain run, launch list, ncu --set full of k_voigt_tile and k_far_nodes
set -x
mkdir -p gpurun_out
T=r2w
summ() {
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/$1_raw.csv > gpurun_out/$1_summary.txt 2>&1
  rm -f gpurun_out/$1.ncu-rep gpurun_out/$1_raw.csv
}
python tools/prof_run.py k1b 16 > gpurun_out/${T}_k1b_plain.log 2>&1; cat gpurun_out/${T}_k1b_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_k1_launches.csv python tools/prof_run.py k1b 16 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_voigt_tile -s 3 -c 1 -o gpurun_out/${T}_tile -f python tools/prof_run.py k1b 16 > gpurun_out/${T}_tile_ncu.log 2>&1; echo "tile rc $?"; summ ${T}_tile
ncu --set full --clock-control none --import-source on -k regex:k_far_nodes -s 3 -c 1 -o gpurun_out/${T}_far -f python tools/prof_run.py k1b 16 > gpurun_out/${T}_far_ncu.log 2>&1; echo "far rc $?"; summ ${T}_far
ls -la gpurun_out

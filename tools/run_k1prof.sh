set -x
mkdir -p gpurun_out
T=${TAG:-p1}
python tools/prof_run.py k1b > gpurun_out/${T}_k1b_plain.log 2>&1; cat gpurun_out/${T}_k1b_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_k1b_launches.csv python tools/prof_run.py k1b > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_voigt_tile -s 3 -c 1 -o gpurun_out/${T}_tile -f python tools/prof_run.py k1b > gpurun_out/${T}_tile_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_core_eval -s 3 -c 1 -o gpurun_out/${T}_core -f python tools/prof_run.py k1b > gpurun_out/${T}_core_ncu.log 2>&1

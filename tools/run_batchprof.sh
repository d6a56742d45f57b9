set -x
mkdir -p gpurun_out
T=${TAG:-p2}
python tools/prof_run.py batch > gpurun_out/${T}_batch_plain.log 2>&1; cat gpurun_out/${T}_batch_plain.log
python tools/prof_run.py batchjac > gpurun_out/${T}_batchjac_plain.log 2>&1; cat gpurun_out/${T}_batchjac_plain.log
SR_PROF_NPIX=500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_batch_launches.csv python tools/prof_run.py batch > /dev/null 2>&1
SR_PROF_NPIX=300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_batchjac_launches.csv python tools/prof_run.py batchjac > /dev/null 2>&1
SR_PROF_NPIX=300 ncu --set full --clock-control none --import-source on -k regex:k_los_layers_jac -s 2 -c 1 -o gpurun_out/${T}_jac -f python tools/prof_run.py batchjac > gpurun_out/${T}_jac_ncu.log 2>&1

# final check of round 2: the new reference-shaped callers first, then the rest of the GPU suite,
# then the bench line if box time is left
mkdir -p gpurun_out
T=${TAG:-r2fin}
NEW="tests/test_gpu_ref_golden2.py tests/test_gpu_latlin.py tests/test_gpu_examples.py"
SECONDS=0
timeout 260 python -m pytest $NEW -m gpu -q --tb=long -rA > gpurun_out/${T}_new.log 2>&1; echo "new rc $? after $SECONDS s"
tail -5 gpurun_out/${T}_new.log
grep -E "^(PASSED|FAILED|ERROR)" gpurun_out/${T}_new.log | head -40
SECONDS=0
timeout 200 python -m pytest tests -m gpu -q --tb=short --ignore=tests/test_gpu_ref_golden2.py --ignore=tests/test_gpu_latlin.py --ignore=tests/test_gpu_examples.py > gpurun_out/${T}_rest.log 2>&1; echo "rest rc $? after $SECONDS s"
tail -4 gpurun_out/${T}_rest.log
grep -E "^FAILED|^ERROR" gpurun_out/${T}_rest.log | head -20
SECONDS=0
timeout 240 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc $? after $SECONDS s"
tail -c 1500 gpurun_out/${T}_bench.json

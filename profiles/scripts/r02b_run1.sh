set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r1_smi.txt 2>&1
timeout 600 python tests/ext_check.py --time > gpurun_out/r1_ext_check.jsonl 2> gpurun_out/r1_ext_check.err; echo "ext_check exit $?" >> gpurun_out/r1_ext_check.err
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r1_pytest.log
timeout 300 python bench.py --only full --steps 3 --warmup 2 > gpurun_out/r1_bench_full.json 2> gpurun_out/r1_bench_full.err
CSIM_FULL_LEGACY=1 timeout 300 python bench.py --only full --steps 3 --warmup 2 > gpurun_out/r1_bench_full_legacy.json 2>> gpurun_out/r1_bench_full.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_trials_dense32 --launch-skip 1 -c 1 -o gpurun_out/r1_prof_ext python profiles/scripts/prof_full.py 250000 > gpurun_out/r1_ncu.log 2>&1
tail -3 gpurun_out/r1_pytest.log; tail -5 gpurun_out/r1_ext_check.jsonl; cat gpurun_out/r1_bench_full.json

mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_probs.py -m gpu -q -k "fast_path" > gpurun_out/r6_pytest_probs.log 2>&1; echo "exit $?" >> gpurun_out/r6_pytest_probs.log
timeout 400 ncu --set full --clock-control none --import-source on --launch-skip 1 -c 1 -k regex:k_trials_prefix -o gpurun_out/r6_prof_prefix python profiles/scripts/prof_pfx.py model 135 1250000 > gpurun_out/r6_ncu_prefix.log 2>&1
tail -n 15 gpurun_out/r6_pytest_probs.log

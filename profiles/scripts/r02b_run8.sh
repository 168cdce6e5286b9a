mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prefix.py tests/test_gpu_probs.py -m gpu -q > gpurun_out/r8_pytest.log 2>&1; echo "exit $?" >> gpurun_out/r8_pytest.log
timeout 200 python profiles/scripts/tune_ext.py > gpurun_out/r8_tune.jsonl 2>&1
tail -n 4 gpurun_out/r8_pytest.log; cat gpurun_out/r8_tune.jsonl

mkdir -p gpurun_out
timeout 65 python -m pytest tests/test_gpu_probs.py tests/test_gpu_prefix.py -m gpu -q > gpurun_out/r13_pytest.log 2>&1; echo "exit $?" >> gpurun_out/r13_pytest.log
tail -n 3 gpurun_out/r13_pytest.log

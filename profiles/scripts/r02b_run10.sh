mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_probs.py -m gpu -q -k "full_size" > gpurun_out/r10_pytest.log 2>&1; echo "exit $?" >> gpurun_out/r10_pytest.log
tail -n 30 gpurun_out/r10_pytest.log

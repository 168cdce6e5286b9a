mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prefix.py -m gpu -q -s -k "term_by_term" > gpurun_out/r7_pytest.log 2>&1; echo "exit $?" >> gpurun_out/r7_pytest.log
tail -n 25 gpurun_out/r7_pytest.log

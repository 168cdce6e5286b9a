mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_prefix.py -m gpu -q -k "full_engine or term_by_term or fall" > gpurun_out/r12_pytest.log 2>&1; echo "exit $?" >> gpurun_out/r12_pytest.log
tail -n 3 gpurun_out/r12_pytest.log

mkdir -p gpurun_out
CSIM_PARITY_REPORT=$PWD/gpurun_out/g_parity_rates.jsonl timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/g_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/g_pytest.log
tail -n 3 gpurun_out/g_pytest.log

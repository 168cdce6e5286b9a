mkdir -p gpurun_out
CSIM_PARITY_REPORT=$PWD/gpurun_out/z_parity_fuzz.jsonl CSIM_FUZZ_CASES=600 timeout 900 python -m pytest tests/test_gpu_prefix.py -m gpu -q -k fuzz > gpurun_out/z_fuzz_600.log 2>&1; echo "exit $?" >> gpurun_out/z_fuzz_600.log
CSIM_PARITY_REPORT=$PWD/gpurun_out/z_parity_fuzz.jsonl CSIM_FUZZ_BIG=1 CSIM_FUZZ_CASES=120 timeout 900 python -m pytest tests/test_gpu_prefix.py -m gpu -q -k fuzz > gpurun_out/z_fuzz_big_120.log 2>&1; echo "exit $?" >> gpurun_out/z_fuzz_big_120.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/z_smoke.log 2>&1; echo "exit $?" >> gpurun_out/z_smoke.log
tail -n 3 gpurun_out/z_fuzz_600.log gpurun_out/z_fuzz_big_120.log gpurun_out/z_smoke.log

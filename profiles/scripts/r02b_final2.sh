# Last run of round 2b on the final tree: GPU suite (with the parity-rate report) and the default bench line
mkdir -p gpurun_out
CSIM_PARITY_REPORT=$PWD/gpurun_out/g_parity_rates.jsonl timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/g_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/g_pytest.log
timeout 900 python bench.py > gpurun_out/g_bench_1gpu.json 2> gpurun_out/g_bench.err; echo "bench exit $?" >> gpurun_out/g_bench.err
tail -n 3 gpurun_out/g_pytest.log; tail -n 1 gpurun_out/g_bench.err; head -c 300 gpurun_out/g_bench_1gpu.json

set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/m_pytest_multi.log 2>&1; echo "exit $?" >> gpurun_out/m_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/m_multi_gpu_check_2gpu.json 2> gpurun_out/m_multi_gpu_check.err; echo "exit $?" >> gpurun_out/m_multi_gpu_check.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/m_bench_2gpu.json 2> gpurun_out/m_bench_2gpu.err; echo "exit $?" >> gpurun_out/m_bench_2gpu.err
tail -3 gpurun_out/m_pytest_multi.log; tail -2 gpurun_out/m_multi_gpu_check.err; head -c 400 gpurun_out/m_multi_gpu_check_2gpu.json; head -c 300 gpurun_out/m_bench_2gpu.json

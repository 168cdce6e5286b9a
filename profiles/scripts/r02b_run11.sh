mkdir -p gpurun_out
timeout 300 python tests/ext_check.py --quick > gpurun_out/r11_ext_check.jsonl 2> gpurun_out/r11_ext_check.err
timeout 900 python -m pytest tests/test_gpu_probs.py tests/test_gpu_prefix.py -m gpu -q > gpurun_out/r11_pytest.log 2>&1; echo "exit $?" >> gpurun_out/r11_pytest.log
timeout 200 python profiles/scripts/tune_ext.py > gpurun_out/r11_tune.jsonl 2>&1
tail -n 3 gpurun_out/r11_pytest.log; cat gpurun_out/r11_tune.jsonl; grep -c '"ext": {"disagree": 0' gpurun_out/r11_ext_check.jsonl

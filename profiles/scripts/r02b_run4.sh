mkdir -p gpurun_out
for lib in ne2_t512 ne1_t512 ne1_t640 ne1_t768; do
CSIM_LIB=$PWD/conclave_sim_b200/_variants/lib_$lib.so timeout 200 python profiles/scripts/tune_ext.py >> gpurun_out/r4_tune.jsonl 2>> gpurun_out/r4_tune.err
done
CSIM_LIB=$PWD/conclave_sim_b200/_variants/lib_ne1_t640.so timeout 300 python tests/ext_check.py --quick > gpurun_out/r4_ext_check_ne1.jsonl 2> gpurun_out/r4_ext_check.err
cat gpurun_out/r4_tune.jsonl; grep -c '"disagree": 0' gpurun_out/r4_ext_check_ne1.jsonl

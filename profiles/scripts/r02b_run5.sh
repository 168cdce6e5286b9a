mkdir -p gpurun_out
for lib in x_ne1_t576 x_ne1_t608 x_ne1_t640; do
CSIM_LIB=$PWD/conclave_sim_b200/_variants/lib_$lib.so timeout 200 python profiles/scripts/tune_ext.py >> gpurun_out/r5_tune.jsonl 2>> gpurun_out/r5_tune.err
done
for lib in default d_ne1_b3 d_ne1_b4 d_ne2_b4 d_ne4_b2; do
echo $lib >> gpurun_out/r5_dense.log
CSIM_LIB=$PWD/conclave_sim_b200/_variants/lib_$lib.so timeout 200 python profiles/scripts/prof_dense.py ref_cli 1250000 >> gpurun_out/r5_dense.log 2>> gpurun_out/r5_tune.err
CSIM_LIB=$PWD/conclave_sim_b200/_variants/lib_$lib.so timeout 200 python profiles/scripts/prof_dense.py model 1250000 >> gpurun_out/r5_dense.log 2>> gpurun_out/r5_tune.err
done
cat gpurun_out/r5_tune.jsonl gpurun_out/r5_dense.log

mkdir -p gpurun_out
timeout 300 python tests/ext_check.py --quick > gpurun_out/r3_ext_check.jsonl 2> gpurun_out/r3_ext_check.err; echo "ext_check exit $?" >> gpurun_out/r3_ext_check.err
for lib in play1 play4; do for wpc in 16 12 8; do
CSIM_LIB=$PWD/conclave_sim_b200/_variants/lib_$lib.so CSIM_EXT_WPC=$wpc timeout 200 python profiles/scripts/tune_ext.py >> gpurun_out/r3_tune.jsonl 2>> gpurun_out/r3_tune.err
done; done
CSIM_FULL_LEGACY=1 timeout 200 python profiles/scripts/tune_ext.py >> gpurun_out/r3_tune.jsonl 2>> gpurun_out/r3_tune.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_trials_dense32 --launch-skip 1 -c 1 -o gpurun_out/r3_prof_ext python profiles/scripts/prof_full.py 250000 > gpurun_out/r3_ncu.log 2>&1
cat gpurun_out/r3_tune.jsonl; grep -c trials gpurun_out/r3_ext_check.jsonl

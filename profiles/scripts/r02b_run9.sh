mkdir -p gpurun_out
for lib in t768 t640; do for wpc in 20 16 12; do
CSIM_LIB=$PWD/conclave_sim_b200/_variants/lib_$lib.so CSIM_EXT_WPC=$wpc timeout 200 python profiles/scripts/tune_ext.py 2>/dev/null | grep "fat+stop" >> gpurun_out/r9_tune.jsonl
done; done
cat gpurun_out/r9_tune.jsonl

set -x
mkdir -p gpurun_out
timeout 600 python tests/ext_check.py --quick --time > gpurun_out/r2_ext_check.jsonl 2> gpurun_out/r2_ext_check.err; echo "ext_check exit $?" >> gpurun_out/r2_ext_check.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_trials_dense32 --launch-skip 1 -c 1 -o gpurun_out/r2_prof_ext python profiles/scripts/prof_full.py 250000 > gpurun_out/r2_ncu.log 2>&1
grep '"time"' gpurun_out/r2_ext_check.jsonl

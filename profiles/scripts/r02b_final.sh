# Final evidence run of round 2b on one B200: GPU suite, ncu captures of the five trial kernels + digest, bench line, launch list, reference arm.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/f_smi.txt 2>&1
CSIM_PARITY_REPORT=$PWD/gpurun_out/f_parity_rates.jsonl timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/f_pytest.log
NCU="ncu --set full --clock-control none --import-source on --launch-skip 1 -c 1"
python profiles/scripts/prof_dense.py ref_cli 1250000 > gpurun_out/f_prof_plain.log 2>&1 && timeout 600 $NCU -k regex:k_trials_dense32 -o gpurun_out/f_prof_dense python profiles/scripts/prof_dense.py ref_cli 1250000 > gpurun_out/f_ncu_dense.log 2>&1
python profiles/scripts/prof_tab.py 1250000 >> gpurun_out/f_prof_plain.log 2>&1 && timeout 600 $NCU -k regex:k_trials_table -o gpurun_out/f_prof_table python profiles/scripts/prof_tab.py 1250000 > gpurun_out/f_ncu_table.log 2>&1
python profiles/scripts/prof_pfx.py model 135 1250000 >> gpurun_out/f_prof_plain.log 2>&1 && timeout 600 $NCU -k regex:k_trials_prefix -o gpurun_out/f_prof_prefix python profiles/scripts/prof_pfx.py model 135 1250000 > gpurun_out/f_ncu_prefix.log 2>&1
python profiles/scripts/prof_full.py 1250000 >> gpurun_out/f_prof_plain.log 2>&1 && timeout 600 $NCU -k regex:k_trials_dense32 -o gpurun_out/f_prof_ext python profiles/scripts/prof_full.py 1250000 > gpurun_out/f_ncu_ext.log 2>&1
CSIM_FULL_LEGACY=1 python profiles/scripts/prof_full.py 250000 >> gpurun_out/f_prof_plain.log 2>&1 && CSIM_FULL_LEGACY=1 timeout 600 $NCU -k regex:k_trials_full32 -o gpurun_out/f_prof_full python profiles/scripts/prof_full.py 250000 > gpurun_out/f_ncu_full.log 2>&1
python profiles/ncu_summary.py --digest profiles/r02b_ncu_digest.json k_trials_dense32=gpurun_out/f_prof_dense.ncu-rep:1250000 k_trials_table=gpurun_out/f_prof_table.ncu-rep:1250000 k_trials_prefix=gpurun_out/f_prof_prefix.ncu-rep:1250000 k_trials_dense32_ext=gpurun_out/f_prof_ext.ncu-rep:1250000 k_trials_full32=gpurun_out/f_prof_full.ncu-rep:250000 > gpurun_out/f_digest.log 2>&1
cp profiles/r02b_ncu_digest.json gpurun_out/f_ncu_digest.json
T="ncu --set full --clock-control none, round 2b final build, cfg-2 parameters, E = 135"
python profiles/ncu_summary.py gpurun_out/f_prof_dense.ncu-rep "$T: profiles/scripts/prof_dense.py ref_cli 1250000" > gpurun_out/f_ncu_dense_summary.txt 2>&1
python profiles/ncu_summary.py gpurun_out/f_prof_table.ncu-rep "$T: profiles/scripts/prof_tab.py 1250000" > gpurun_out/f_ncu_table_summary.txt 2>&1
python profiles/ncu_summary.py gpurun_out/f_prof_prefix.ncu-rep "$T: profiles/scripts/prof_pfx.py model 135 1250000" > gpurun_out/f_ncu_prefix_model_summary.txt 2>&1
python profiles/ncu_summary.py gpurun_out/f_prof_ext.ncu-rep "$T + candidate fatigue + stop-candidate: profiles/scripts/prof_full.py 1250000 (FULL engine fast path)" > gpurun_out/f_ncu_ext_summary.txt 2>&1
python profiles/ncu_summary.py gpurun_out/f_prof_full.ncu-rep "$T + candidate fatigue + stop-candidate: CSIM_FULL_LEGACY=1 profiles/scripts/prof_full.py 250000 (k_trials_full32)" > gpurun_out/f_ncu_full_legacy_summary.txt 2>&1
rm -f gpurun_out/f_prof_table.ncu-rep gpurun_out/f_prof_prefix.ncu-rep gpurun_out/f_prof_full.ncu-rep
timeout 900 python bench.py > gpurun_out/f_bench_1gpu.json 2> gpurun_out/f_bench.err; echo "bench exit $?" >> gpurun_out/f_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_ncu_launches.csv python bench.py --steps 2 --warmup 3 --trials-per-gpu 250000 --no-cpu-baseline --no-configs > gpurun_out/f_ncu_launches.log 2>&1
timeout 600 python bench.py --impl reference > gpurun_out/f_bench_reference.json 2> gpurun_out/f_bench_reference.err
tail -3 gpurun_out/f_pytest.log; cat gpurun_out/f_prof_plain.log; head -c 600 gpurun_out/f_bench_1gpu.json; ls -la gpurun_out

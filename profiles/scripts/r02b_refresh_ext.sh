# refresh the FULL fast path's evidence after a kernel change: plain run, ncu capture, digest entry, summary, bench line
mkdir -p gpurun_out
python profiles/scripts/prof_full.py 1250000 > gpurun_out/h_prof_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 1 -c 1 -k regex:k_trials_dense32 -o gpurun_out/h_prof_ext python profiles/scripts/prof_full.py 1250000 > gpurun_out/h_ncu_ext.log 2>&1
python profiles/ncu_summary.py --digest gpurun_out/h_digest_ext.json k_trials_dense32_ext=gpurun_out/h_prof_ext.ncu-rep:1250000 > /dev/null 2>&1
python - <<'PY'
import json
d = json.load(open('profiles/r02b_ncu_digest.json')); e = json.load(open('gpurun_out/h_digest_ext.json'))
d.update(e); json.dump(d, open('profiles/r02b_ncu_digest.json', 'w'), indent=1); json.dump(d, open('gpurun_out/h_ncu_digest.json', 'w'), indent=1)
PY
python profiles/ncu_summary.py gpurun_out/h_prof_ext.ncu-rep "ncu --set full --clock-control none, round 2b final build, cfg-2 parameters, E = 135 + candidate fatigue + stop-candidate: profiles/scripts/prof_full.py 1250000 (FULL engine fast path)" > gpurun_out/h_ncu_ext_summary.txt 2>&1
timeout 900 python bench.py > gpurun_out/h_bench_1gpu.json 2> gpurun_out/h_bench.err; echo "bench exit $?" >> gpurun_out/h_bench.err
cat gpurun_out/h_prof_plain.log; tail -n 1 gpurun_out/h_bench.err

mkdir -p gpurun_out
which compute-sanitizer > gpurun_out/s_which.txt 2>&1
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/ext_check.py --quick > gpurun_out/s_memcheck.log 2>&1; echo "exit $?" >> gpurun_out/s_memcheck.log
timeout 240 compute-sanitizer --tool racecheck --error-exitcode 9 python tests/ext_check.py --quick > gpurun_out/s_racecheck.log 2>&1; echo "exit $?" >> gpurun_out/s_racecheck.log
tail -n 8 gpurun_out/s_memcheck.log; tail -n 8 gpurun_out/s_racecheck.log

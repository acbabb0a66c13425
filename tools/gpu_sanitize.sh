set -x
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_target.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitize_target.py > gpurun_out/r02_sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -4 gpurun_out/r02_sanitizer_synccheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_target.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/r02_sanitizer_racecheck.log

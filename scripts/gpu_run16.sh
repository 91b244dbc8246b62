mkdir -p gpurun_out/r2
echo "=== memcheck"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/r2/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/r2/sanitizer_memcheck.log
echo "=== racecheck"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/r2/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -4 gpurun_out/r2/sanitizer_racecheck.log
echo "=== initcheck"
timeout 900 compute-sanitizer --tool initcheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/r2/sanitizer_initcheck.log 2>&1; echo "initcheck rc=$?"
tail -4 gpurun_out/r2/sanitizer_initcheck.log

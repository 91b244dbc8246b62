# final single-GPU check of the round: the GPU suite and smoke() on the final code
mkdir -p gpurun_out/r2
timeout 420 python -m pytest tests -x -q -m gpu > gpurun_out/r2/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2/pytest_gpu_final.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2

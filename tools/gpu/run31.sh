set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final4.log 2>&1; echo "pytest final rc=$?"; tail -6 gpurun_out/pytest_gpu_final4.log

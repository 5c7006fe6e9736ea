set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_cholesky.py tests/test_gpu_cpp_mirror.py tests/test_gpu_device_matrix.py -q 2>&1 | tail -4
timeout 120 python tools/chol_profile.py 16384 3 2>&1 | tail -3

set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_lu_mg.py -q -k "python_mirror" 2>&1 | tail -6

set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_lu_mg.py -q -k "not 32768 and not 16384" 2>&1 | tail -3

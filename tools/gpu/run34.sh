set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_lu_mg.py -q > gpurun_out/pytest_lu_mg_nb.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_lu_mg_nb.log
timeout 100 python tools/lu_mg_profile.py 32768 1 0 0,0 2>&1 | tail -2

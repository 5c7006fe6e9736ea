set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lu_mg.py -x -q > gpurun_out/pytest_lu_mg.log 2>&1; echo "pytest lu_mg rc=$?"; tail -15 gpurun_out/pytest_lu_mg.log
timeout 300 python tools/lu_mg_profile.py 16384 3 0 0,0 0,0,0,0 2>&1 | tail -5
timeout 200 python tools/lu_mg_profile.py 8192 3 0 0,0 2>&1 | tail -3

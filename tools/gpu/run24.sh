set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_lu_mg.py -x -q -k "distinct" > gpurun_out/pytest_lu_mg4.log 2>&1; echo "pytest lu_mg rc=$?"; tail -6 gpurun_out/pytest_lu_mg4.log
timeout 300 python tools/lu_mg_profile.py 16384 3 0 0,1 0,1,2,3 2>&1 | tail -5
timeout 300 python tools/lu_mg_profile.py 28672 2 0 0,1 0,1,2,3 2>&1 | tail -5

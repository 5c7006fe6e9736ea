set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m pytest tests/test_gpu_lu_mg.py -x -q -k "distinct_devices_matches_fixture and 8" > gpurun_out/pytest_lu_mg8.log 2>&1; echo "pytest lu_mg rc=$?"; tail -4 gpurun_out/pytest_lu_mg8.log
LA_LU_MG_TRACE=1 timeout 200 python tools/lu_mg_profile.py 16384 3 0,1,2,3,4,5,6,7 --check 2>&1 | tail -5
LA_LU_MG_TRACE=1 timeout 200 python tools/lu_mg_profile.py 28672 2 0,1,2,3,4,5,6,7 2>&1 | tail -4

set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 100 python tools/lu_mg_profile.py 32768 2 0,1 2>&1 | tail -1

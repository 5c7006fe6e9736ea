set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for d in 32 48 64 96 128; do LA_LU_RPC_DIV=$d timeout 60 python tools/lu_profile.py 16384 4 2>&1 | tail -2 | sed "s/^/rpc_div=$d /"; done

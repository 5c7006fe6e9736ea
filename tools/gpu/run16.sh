set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for d in 48 24 32 64 96 148; do echo "== LA_QR_RPC_DIV=$d"; LA_QR_RPC_DIV=$d timeout 100 python tools/qr_profile.py 16384 16384 2 | tail -1; done

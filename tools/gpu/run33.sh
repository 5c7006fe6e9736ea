set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 200 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_small.py gemm32 lu_mg chol > gpurun_out/sanitize_memcheck3.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/sanitize_memcheck3.log
timeout 200 compute-sanitizer --tool synccheck --print-limit 5 python tools/sanitize_small.py gemm32 lu_mg chol > gpurun_out/sanitize_synccheck3.log 2>&1; echo "synccheck rc=$?"; tail -5 gpurun_out/sanitize_synccheck3.log

set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_lu_parity.py tests/test_gpu_cholesky.py tests/test_gpu_full_size.py tests/test_gpu_reference_tests.py tests/test_gpu_config0.py -q -k "solve or chol or 16384 or block_rows or config" 2>&1 | tail -4
timeout 100 python tools/lu_profile.py 16384 4 --solve
LA_SOLVE_COMBINE=0 timeout 100 python tools/lu_profile.py 16384 3 --solve
LA_SOLVE_CHAINS=1 timeout 100 python tools/lu_profile.py 16384 3 --solve
LA_SOLVE_TRACE=1 timeout 100 python tools/lu_profile.py 16384 1 --solve 2>&1 | grep -E "sweep2 fwd" | head -8

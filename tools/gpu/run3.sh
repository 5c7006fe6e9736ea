set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
LA_QR_LOOKAHEAD=0 timeout 120 python tools/gpu/qr_debug.py > gpurun_out/qr_debug.log 2>&1; tail -60 gpurun_out/qr_debug.log
timeout 120 python tools/gpu/qr_debug.py 2>&1 | grep -E "shape|err" | tail -12
LA_QR_LOOKAHEAD=0 timeout 300 python -m pytest tests/test_gpu_qr.py -q 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_qr.py tests/test_gpu_elementwise.py tests/test_gpu_cpp_mirror.py -q > gpurun_out/pytest_qr.log 2>&1; echo "pytest qr rc=$?"; tail -40 gpurun_out/pytest_qr.log
for i in 1 2 3; do
  LA_MG_CHILD=1 LA_MG_TRACE=1 timeout 100 python -m pytest tests/test_gpu_mg.py -x -q -s -k "one_device_host" > gpurun_out/mg_trace_$i.log 2>&1; echo "mg trace run $i rc=$?"
done
for i in 1 2; do
  LA_MG_CHILD=1 LA_MG_TRACE=1 CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 100 python -m pytest tests/test_gpu_mg.py -x -q -s -k "one_device_host" > gpurun_out/mg_trace32_$i.log 2>&1; echo "mg trace32 run $i rc=$?"
done
tail -30 gpurun_out/mg_trace32_1.log
timeout 200 python tools/qr_profile.py 4096 4096 2
timeout 200 python tools/qr_profile.py 16384 16384 2
LA_QR_LOOKAHEAD=0 timeout 200 python tools/qr_profile.py 16384 16384 2
LA_LU_TRACE=gpurun_out/lu_trace_r2.csv timeout 100 python tools/lu_profile.py 16384 2
timeout 300 python -m pytest tests/test_gpu_lu_parity.py tests/test_gpu_cholesky.py tests/test_gpu_full_size.py -q -k "solve or chol or 16384" > gpurun_out/pytest_solve.log 2>&1; echo "pytest solve rc=$?"; tail -15 gpurun_out/pytest_solve.log
timeout 100 python tools/lu_profile.py 16384 3 --solve
LA_SOLVE_OLD_SWEEP=1 timeout 100 python tools/lu_profile.py 16384 2 --solve

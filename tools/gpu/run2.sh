set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --durations=12 -k "not mg" > gpurun_out/pytest_nomg.log 2>&1; echo "pytest(not mg) rc=$?"; tail -25 gpurun_out/pytest_nomg.log
timeout 300 python -X faulthandler -m pytest tests/test_gpu_mg.py -x -q -k "one_device" -o faulthandler_timeout=50 > gpurun_out/pytest_mg1.log 2>&1; echo "mg one-device rc=$?"; tail -60 gpurun_out/pytest_mg1.log
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 python -X faulthandler -m pytest tests/test_gpu_mg.py -x -q -k "one_device" -o faulthandler_timeout=50 > gpurun_out/pytest_mg2.log 2>&1; echo "mg one-device (32 connections) rc=$?"; tail -30 gpurun_out/pytest_mg2.log
for g in 1 2 3; do echo "LA_LU_GROUP=$g"; LA_LU_GROUP=$g timeout 120 python tools/lu_profile.py 16384 4; done
LA_LU_GROUP=2 LA_LU_GROUP_ROWS=6144 timeout 120 python tools/lu_profile.py 16384 3
LA_LU_GROUP=2 LA_LU_GROUP_ROWS=1024 timeout 120 python tools/lu_profile.py 16384 3
timeout 120 python tools/gemm_bench.py 16256,128,16256,1 16256,256,16256,1 16256,384,16256,1 8192,128,8192,1 8192,256,8192,1 128,16384,16256,0 128,8192,8192,0
timeout 200 python tools/qr_profile.py 4096 4096 2
timeout 200 python tools/qr_profile.py 16384 16384 2

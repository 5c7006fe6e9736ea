set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
LA_TF32_PAIR=1 LA_GEMM_F32_MODE=tf32 timeout 60 python tools/gemm_bench_f32.py 512,256,512,2 2>&1 | tail -2; echo "small rc=$?"
LA_TF32_PAIR=1 LA_GEMM_F32_MODE=tf32 timeout 90 python tools/gemm_bench_f32.py 4096,1024,4096,2 65536,1024,16384,2 8192,8192,8192,2 2>&1 | tail -3; echo "pair rc=$?"
LA_TF32_PAIR=0 LA_GEMM_F32_MODE=tf32 timeout 90 python tools/gemm_bench_f32.py 4096,1024,4096,2 65536,1024,16384,2 8192,8192,8192,2 2>&1 | tail -3
LA_TF32_PAIR=1 timeout 90 python tools/gemm_bench_f32.py 65536,1024,16384,2 2>&1 | tail -1
LA_TF32_PAIR=0 timeout 90 python tools/gemm_bench_f32.py 65536,1024,16384,2 2>&1 | tail -1
LA_TF32_PAIR=1 timeout 300 python -m pytest tests/test_gpu_gemm_parity.py -x -q -k f32 2>&1 | tail -5

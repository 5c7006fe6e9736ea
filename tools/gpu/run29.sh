set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
LA_GEMM_F32_MODE=tf32 timeout 300 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 2 -c 1 -o gpurun_out/r2_gemm_f32_pair -f python tools/gemm_bench_f32.py 65536,1024,16384,2 > gpurun_out/ncu_pair.log 2>&1; echo "ncu pair rc=$?"; tail -2 gpurun_out/ncu_pair.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 1 --skip-cpu > gpurun_out/ncu_bench.log 2>&1; echo "ncu bench rc=$?"; tail -c 600 gpurun_out/ncu_bench.log
LA_GEMM_F32_MODE=tf32 timeout 90 python tools/gemm_bench_f32.py 65536,1024,16384,2 8192,8192,8192,2 16384,16384,16384,2 2>&1 | tail -3
timeout 90 python tools/gemm_bench_f32.py 65536,1024,16384,2 2>&1 | tail -1

set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for g in 4 8 16 32 64 128; do LA_TF32_GROUP_M=$g LA_GEMM_F32_MODE=tf32 timeout 90 python tools/gemm_bench_f32.py 65536,1024,16384,2 8192,8192,8192,2 2>&1 | tail -2 | sed "s/^/group_m=$g /"; done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final2.log 2>&1; echo "pytest final rc=$?"; tail -8 gpurun_out/pytest_gpu_final2.log
timeout 600 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_final2.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2

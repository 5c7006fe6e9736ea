# full GPU test suite + bench + launch list + sanitizer (1 GPU)
set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest full rc=$?"; tail -30 gpurun_out/pytest_gpu_full.log
timeout 600 python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo "bench rc=$?"; tail -c 6000 gpurun_out/bench_r2.json; tail -5 gpurun_out/bench_r2.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for tool in racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py gemm64 gemm32 lu solve chol qr > gpurun_out/sanitize_$tool.log 2>&1; echo "sanitizer $tool rc=$?"; tail -8 gpurun_out/sanitize_$tool.log
done

set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
LA_CHOL_LOWER=0 timeout 120 python tools/chol_profile.py 16384 3 2>&1 | tail -3
LA_CHOL_LOWER=1 timeout 120 python tools/chol_profile.py 16384 3 2>&1 | tail -3
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final3.log 2>&1; echo "pytest final rc=$?"; tail -6 gpurun_out/pytest_gpu_final3.log
timeout 400 python bench.py > gpurun_out/bench_final3.json 2> gpurun_out/bench_final3.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_final3.err
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2

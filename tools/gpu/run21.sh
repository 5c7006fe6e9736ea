set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest final rc=$?"; tail -8 gpurun_out/pytest_gpu_final.log
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_final.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_lu16384_solve_launches.csv python tools/lu_profile.py 16384 1 --solve > gpurun_out/ncu_lu.log 2>&1; echo "ncu lu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_sweep_mma -c 2 -o gpurun_out/r2_solve_sweep -f python tools/lu_profile.py 16384 1 --solve > gpurun_out/ncu_sweep.log 2>&1; echo "ncu sweep rc=$?"

set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for i in 1 2 3; do timeout 400 python -m pytest tests/test_gpu_mg.py -q -k "one_device" 2>&1 | tail -4; done
timeout 300 python -m pytest tests/test_gpu_qr.py tests/test_gpu_lu_parity.py -q 2>&1 | tail -4
LA_SOLVE_TRACE=1 timeout 100 python tools/lu_profile.py 16384 1 --solve 2>&1 | tail -80
LA_SOLVE_TRACE=1 LA_SOLVE_CHAINS=1 timeout 100 python tools/lu_profile.py 16384 1 --solve 2>&1 | grep -E "sweep2 fwd|solve" | tail -40
timeout 100 python tools/lu_profile.py 16384 3 --solve
LA_SOLVE_CHAINS=1 timeout 100 python tools/lu_profile.py 16384 3 --solve
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python tools/sanitize_small.py lu solve qr > gpurun_out/sanitize_racecheck2.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/sanitize_racecheck2.log
bash tools/gpu/run5.sh

# profiles: launch list of one LU + solve + QR, full captures of the dominant kernels (1 GPU)
set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_lu16384_solve_launches.csv python tools/lu_profile.py 16384 1 --solve > gpurun_out/ncu_lu.log 2>&1; echo "ncu lu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_qr8192_launches.csv python tools/qr_profile.py 8192 8192 1 > gpurun_out/ncu_qr.log 2>&1; echo "ncu qr rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_sweep_mma -c 2 -o gpurun_out/r2_solve_sweep -f python tools/lu_profile.py 16384 1 --solve > gpurun_out/ncu_sweep.log 2>&1; echo "ncu sweep rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qr_panel -s 2 -c 1 -o gpurun_out/r2_qr_panel -f python tools/qr_profile.py 8192 8192 1 > gpurun_out/ncu_qrpanel.log 2>&1; echo "ncu qr panel rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f64_tma -s 40 -c 1 -o gpurun_out/r2_gemm_k256_inlu -f python tools/lu_profile.py 16384 1 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
ls -la gpurun_out/*.ncu-rep

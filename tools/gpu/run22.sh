set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lu_panel -s 60 -c 2 -o gpurun_out/r2_lu_panel -f python tools/lu_profile.py 16384 1 > gpurun_out/ncu_lupanel.log 2>&1; echo "ncu lu panel rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f64_tma -s 60 -c 10 -o gpurun_out/r2_gemm_inlu -f python tools/lu_profile.py 16384 1 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chol_diag -s 10 -c 1 -o gpurun_out/r2_chol_diag -f python tools/chol_profile.py 8192 1 > gpurun_out/ncu_chol.log 2>&1; echo "ncu chol rc=$?"
ls -la gpurun_out/*.ncu-rep

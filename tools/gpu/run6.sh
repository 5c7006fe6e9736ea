# 2-GPU: multi-GPU Mul tests (single process two devices, one process per device over IPC) + bench --gpus 2
set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_mg.py -q > gpurun_out/pytest_mg_2gpu.log 2>&1; echo "pytest mg rc=$?"; tail -30 gpurun_out/pytest_mg_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"; tail -c 3000 gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err

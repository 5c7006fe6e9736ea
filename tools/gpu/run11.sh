# 8-GPU scaling check: bench --gpus 8 (f64 32768^3 sharded + f32 config), then mg tests on 8 devices visible
set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "bench8 rc=$?"; tail -c 1500 gpurun_out/bench_8gpu.json; tail -5 gpurun_out/bench_8gpu.err

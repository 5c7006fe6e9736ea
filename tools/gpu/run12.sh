# 2-GPU regression of the fused pull + sweep timing (1 GPU part runs on device 0)
set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mg.py -q 2>&1 | tail -3
timeout 100 python tools/lu_profile.py 16384 4 --solve
timeout 200 python -m pytest tests/test_gpu_lu_parity.py tests/test_gpu_full_size.py -q -k "solve or block_rows" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu_b.json 2> gpurun_out/bench_2gpu_b.err; echo "bench2 rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/bench_2gpu_b.json').read().strip().splitlines()[-1]);print(d['value'],d['e2e']['value'],d['f32']['tflops'],d['f32']['ms'])"

set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err; echo "bench4 rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_4gpu.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],d['f32']['tflops'],d['f32']['ms'],d['parity']['ok'])"
timeout 300 python bench.py --impl reference --gpus 4 --steps 2 --warmup 1 | cut -c1-300

set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu_d.json 2> gpurun_out/bench_2gpu_d.err; echo "bench2 rc=$?"
tail -3 gpurun_out/bench_2gpu_d.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_2gpu_d.json') if x.startswith('{')]
d=json.loads(l[-1]); print(d['value'], d['e2e']['value'], json.dumps(d.get('lu_mg'))[:900])
PY

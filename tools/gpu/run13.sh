# 8 GPUs: SM pull (default) vs copy-engine pull for the f64 32768^3 and f32 configs
set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for mode in sm ce; do
  LA_MG_PULL=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_8gpu_$mode.json 2> gpurun_out/bench_8gpu_$mode.err; echo "bench8 $mode rc=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/bench_8gpu_$mode.json').read().strip().splitlines()[-1]);print('$mode',d['value'],d['ms_per_step'],d['e2e']['value'],d['f32']['tflops'],d['f32']['ms'],d['f32']['default_mode_3xtf32']['tflops'])"
done

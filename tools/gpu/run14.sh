set -x
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 300 python tools/gpu/illcond.py
( time timeout 900 python bench.py > gpurun_out/bench_time.json 2>/dev/null ) 2>&1 | tail -4

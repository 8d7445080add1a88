set -x
mkdir -p gpurun_out
G=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $G --steps 20 --warmup 5 2> gpurun_out/r02_bench_${G}gpu.err | grep '^{' > gpurun_out/r02_bench_${G}gpu_lv.json
tail -3 gpurun_out/r02_bench_${G}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_${G}gpu_lv.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','strong','allreduce_check','kernel_ms')}); print(d['e2e'])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29556 bench.py --config hjb --gpus $G --steps 20 --warmup 3 2> gpurun_out/r02_bench_${G}gpu_hjb.err | grep '^{' > gpurun_out/r02_bench_${G}gpu_hjb.json
tail -3 gpurun_out/r02_bench_${G}gpu_hjb.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_${G}gpu_hjb.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e')})
PY

set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests -q -m gpu -k "two_ranks or two_devices or tanh or adaptive_tensor_core" -s 2>&1 | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; tail -3 gpurun_out/r02_bench_2gpu.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_2gpu.json'))
print({k:d[k] for k in ('value','ms_per_step','strong','allreduce_check','kernels_per_step','gpu_launches')})
print(d['e2e'], d['cpu_baseline'])
PY

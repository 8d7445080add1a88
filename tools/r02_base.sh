set -x
mkdir -p gpurun_out
for n in 65536 32768 16384 8192; do
  timeout 300 python bench.py --n-per-gpu $n --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r02_base_n$n.json 2> gpurun_out/r02_base_n$n.err
  tail -c 600 gpurun_out/r02_base_n$n.json
done
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; lscpu | head -20

set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -s 2>&1 | grep -E "SEIR:|device tanh|passed|failed|FAILED|Error|error" | tail -20 > gpurun_out/r02_pytest.txt; cat gpurun_out/r02_pytest.txt
timeout 300 python bench.py --steps 20 --warmup 5 2> gpurun_out/r02_bench_lv.err | grep '^{' > gpurun_out/r02_bench_lv.json
timeout 600 python bench.py --config seir --steps 5 --warmup 3 2> gpurun_out/r02_bench_seir.err | grep '^{' > gpurun_out/r02_bench_seir.json
timeout 600 python bench.py --config fkpp --steps 5 --warmup 3 2> gpurun_out/r02_bench_fkpp.err | grep '^{' > gpurun_out/r02_bench_fkpp.json
python - <<'PY'
import json
for f in ("lv","seir","fkpp"):
    try:
        d=json.loads(open(f"gpurun_out/r02_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), d.get("ms_per_step"), d.get("kernel_ms"), (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("cores"))
    except Exception as e:
        print(f, "ERR", e)
PY

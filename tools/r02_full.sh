set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/r02_pytest.txt; cat gpurun_out/r02_pytest.txt
timeout 300 python bench.py --steps 20 --warmup 5 2> gpurun_out/r02_bench_lv.err | grep '^{' > gpurun_out/r02_bench_lv.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | grep '^{' > gpurun_out/r02_bench_lv_ref.json
timeout 600 python bench.py --config seir --steps 5 --warmup 3 2> gpurun_out/r02_bench_seir.err | grep '^{' > gpurun_out/r02_bench_seir.json
timeout 600 python bench.py --config fkpp --steps 5 --warmup 3 2> gpurun_out/r02_bench_fkpp.err | grep '^{' > gpurun_out/r02_bench_fkpp.json
tail -3 gpurun_out/r02_bench_seir.err gpurun_out/r02_bench_fkpp.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 30 --csv --log-file gpurun_out/r02_ncu_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-strong > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'forward_kernel|adjoint_kernel' -s 4 -c 2 -f -o gpurun_out/r02_wm_65536 python tools/prof_step.py 65536 4 > gpurun_out/r02_prof.log 2>&1
python - <<'PY'
import json
for f in ("lv","lv_ref","seir","fkpp"):
    try:
        d=json.loads(open(f"gpurun_out/r02_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), d.get("ms_per_step"), d.get("kernel_ms"), (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("cores"))
    except Exception as e:
        print(f, "ERR", e)
PY

set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/r02_pytest.txt; tail -40 gpurun_out/r02_pytest.txt

set -x
timeout 120 python tools/seir_dbg.py 2>&1 | tail -14
bash tools/r02_run8.sh 4

set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'forward_kernel|adjoint_kernel' -s 4 -c 2 -f -o gpurun_out/r02_wm_65536 python tools/prof_step.py 65536 4 > gpurun_out/r02_prof.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'forward_kernel|adjoint_kernel' -s 4 -c 2 -f -o gpurun_out/r02_wm_8192 python tools/prof_step.py 8192 4 >> gpurun_out/r02_prof.log 2>&1
tail -5 gpurun_out/r02_prof.log
ls -la gpurun_out/*.ncu-rep

mkdir -p gpurun_out/r02_vb4
TAGS=30000000 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:load_tile_kernel -s 14 -c 1 -f -o gpurun_out/r02_vb4/tile python tools/e2e_trace.py > gpurun_out/r02_vb4/ncu.log 2>&1; echo rc=$?
tail -3 gpurun_out/r02_vb4/ncu.log
ls -la gpurun_out/r02_vb4

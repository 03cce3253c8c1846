mkdir -p gpurun_out/r02_vb2
timeout 600 python -m pytest tests -x -q -m gpu -k "vbyte or packed_load or async_load or cli" > gpurun_out/r02_vb2/pytest_sel.txt 2>&1; echo sel rc=$?
tail -5 gpurun_out/r02_vb2/pytest_sel.txt
UPB_TRACE=1 timeout 300 python tools/e2e_trace.py > gpurun_out/r02_vb2/trace.txt 2>&1; tail -45 gpurun_out/r02_vb2/trace.txt
bash tools/gpu_session.sh r02_vb2 bench-quick

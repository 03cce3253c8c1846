mkdir -p gpurun_out/r02_vb3
timeout 600 python -m pytest tests -x -q -m gpu -k "vbyte or packed_load or async_load" > gpurun_out/r02_vb3/pytest_sel.txt 2>&1; echo sel rc=$?
tail -3 gpurun_out/r02_vb3/pytest_sel.txt
UPB_TRACE=1 timeout 300 python tools/e2e_trace.py > gpurun_out/r02_vb3/trace.txt 2>&1; grep -v "copied at  -1" gpurun_out/r02_vb3/trace.txt | tail -32

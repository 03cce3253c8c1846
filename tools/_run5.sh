mkdir -p gpurun_out/r02_vb5
timeout 600 python -m pytest tests -x -q -m gpu -k "vbyte or packed_load or async_load" > gpurun_out/r02_vb5/pytest_sel.txt 2>&1; echo sel rc=$?
tail -2 gpurun_out/r02_vb5/pytest_sel.txt
for cum in "" "0.08,0.26,0.48,0.66,0.80,0.90,0.96" "0.05,0.15,0.32,0.50,0.66,0.80,0.90,0.96" "0.08,0.24,0.44,0.62,0.78,0.90"; do
  echo "== cum=$cum"
  UPB_LOAD_CUM=$cum timeout 300 python tools/e2e_trace.py 2>&1 | grep -E "load \+ call|load \+ wait|units 46|group 13 .*copied at +[0-9]" | tail -4
done

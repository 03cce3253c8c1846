mkdir -p gpurun_out/r02_fuzz
UPB_FUZZ_SEEDS=600 UPB_FUZZ_SHIFT_SEEDS=96 timeout 900 python -m pytest tests/test_zz_fuzz_gpu.py -q -m gpu -x > gpurun_out/r02_fuzz/pytest_fuzz600.txt 2>&1; echo fuzz rc=$?
tail -8 gpurun_out/r02_fuzz/pytest_fuzz600.txt

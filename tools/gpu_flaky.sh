#!/bin/bash
# is tests/test_host_cli.py::test_cli_regions_nondirectional_shifts_controls_coeffs stable?  current library x3, then an older build x2
set +e
mkdir -p gpurun_out/flaky
for i in 1 2 3; do python -m pytest tests/test_host_cli.py -q -m gpu -k "nondirectional" 2>&1 | tail -1; done
for i in 1 2; do LD_LIBRARY_PATH=$PWD/build_exp/v29 python -m pytest tests/test_host_cli.py -q -m gpu -k "nondirectional" 2>&1 | tail -1; done

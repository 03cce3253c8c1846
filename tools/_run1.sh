mkdir -p gpurun_out/r02_lanes1
timeout 300 python tools/lanes_ab.py > gpurun_out/r02_lanes1/lanes.jsonl 2> gpurun_out/r02_lanes1/lanes.err; echo lanes rc=$?
cat gpurun_out/r02_lanes1/lanes.jsonl | cut -c1-400; tail -5 gpurun_out/r02_lanes1/lanes.err
bash tools/gpu_session.sh r02_lanes1 tests-fast bench-quick

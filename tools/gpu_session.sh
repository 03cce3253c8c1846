#!/bin/bash
# One gpurun call of round 2: smoke, the GPU test suite, the bench line, the ncu launch list and
# one full capture of span_kernel.  Everything lands in gpurun_out/$TAG/.  Steps are chosen by name:
#   bash tools/gpu_session.sh <tag> [smoke] [tests] [tests-fast] [bench] [bench-quick] [onelane] [trace] [ncu] [ncufull] [ncutile] [refarm] [cli] [ab:<workloads>]
set +e
TAG=${1:-r02}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
note () { echo "$1 rc=$2 t=$(( $(date +%s) - T0 ))" >> $OUT/progress.txt; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,temperature.gpu,memory.used --format=csv > $OUT/smi.txt 2>&1
for step in "$@"; do
  case $step in
    smoke)
      timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.txt 2>&1; note smoke $? ;;
    tests)
      timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.txt 2>&1; note pytest $? ;;
    tests-fast)
      timeout 900 python -m pytest tests -x -q -m gpu -k "not full_size and not at_full_size" > $OUT/pytest_gpu_fast.txt 2>&1; note pytest-fast $? ;;
    tests-all)
      timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu_all.txt 2>&1; note pytest-all $? ;;
    bench)
      timeout 600 python bench.py > $OUT/bench_c3.json 2> $OUT/bench_c3.err; note bench $? ;;
    bench-quick)
      timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-sweep > $OUT/bench_quick.json 2> $OUT/bench_quick.err; note bench-quick $? ;;
    ncu)
      timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:upb:: --csv --log-file $OUT/launches.csv \
        python bench.py --steps 2 --warmup 1 --no-cpu --no-sweep > $OUT/ncu_list.log 2>&1; note ncu-list $? ;;
    ncufull)
      timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:upb::span_kernel -c 1 -f -o $OUT/span \
        python bench.py --steps 2 --warmup 1 --no-cpu --no-sweep > $OUT/ncu_full.log 2>&1; note ncu-full $?
      timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:upb::region_pool_kernel -c 1 -f -o $OUT/pool \
        python bench.py --steps 2 --warmup 1 --no-cpu --no-sweep > $OUT/ncu_full_pool.log 2>&1; note ncu-full-pool $? ;;
    ncutile)
      timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:upb::load_tile_kernel -s 14 -c 1 -f -o $OUT/tile \
        python tools/e2e_trace.py > $OUT/ncu_full_tile.log 2>&1; note ncu-full-tile $? ;;
    onelane)
      UPB_CALL_GROUPS=1 timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-sweep > $OUT/bench_quick_onelane.json 2> $OUT/bench_quick_onelane.err; note bench-quick-onelane $? ;;
    trace)
      UPB_TRACE=1 timeout 300 python tools/e2e_trace.py > $OUT/e2e_trace.txt 2>&1; note trace $? ;;
    cli)
      timeout 600 python tools/cli_bench.py > $OUT/cli_bench.jsonl 2> $OUT/cli_bench.err; note cli $? ;;
    wig)
      timeout 600 python tools/wig_bench.py > $OUT/wig_bench.json 2> $OUT/wig_bench.err; note wig $? ;;
    refarm)
      timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > $OUT/reference_arm.json 2> $OUT/reference_arm.err; note refarm $? ;;
    ab:*)
      WL=${step#ab:}
      for so in build_exp/*.so; do
        v=$(basename $so .so)
        UPB_LIB=$PWD/$so timeout 300 python tools/ab.py ${WL//,/ } > $OUT/ab_$v.jsonl 2> $OUT/ab_$v.err; note "ab $v" $?
      done ;;
    *) echo "unknown step $step" >> $OUT/progress.txt ;;
  esac
done
cat $OUT/progress.txt
for f in smoke.txt pytest_gpu.txt pytest_gpu_fast.txt pytest_gpu_all.txt; do [ -f $OUT/$f ] && { echo "== $f"; tail -25 $OUT/$f; }; done
for f in cli_bench.jsonl wig_bench.json reference_arm.json; do [ -f $OUT/$f ] && { echo "== $f"; cut -c1-1500 $OUT/$f; }; done
for f in bench_c3 bench_quick bench_quick_onelane; do [ -f $OUT/$f.json ] && { echo "== $f"; cut -c1-1500 $OUT/$f.json; tail -5 $OUT/$f.err; }; done
exit 0

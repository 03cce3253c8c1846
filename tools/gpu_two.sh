#!/bin/bash
# One gpurun --gpus N call: what needs more than one GPU.  bash tools/gpu_two.sh <tag> <N> [steps...]
set +e
TAG=$1; N=$2; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
T0=$(date +%s)
note () { echo "$1 rc=$2 t=$(( $(date +%s) - T0 ))" >> $OUT/progress.txt; }
nvidia-smi --query-gpu=index,name --format=csv > $OUT/smi.txt 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
for step in "$@"; do
  case $step in
    tests)
      timeout 900 python -m pytest tests/test_zz_multi_gpu.py tests/test_zz_two_gpus.py tests/test_host_cli.py -q -m gpu -k "two_gpus or two-gpus or sharded_over_nccl" > $OUT/pytest_two.txt 2>&1; note pytest-two $? ;;
    shift)
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 tools/shift_sharded_bench.py > $OUT/shift_n$N.json 2> $OUT/shift_n$N.err; note shift $? ;;
    tir)
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29702 tools/tir_sharded_bench.py > $OUT/tir_n$N.json 2> $OUT/tir_n$N.err; note tir $? ;;
    bench)
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29703 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; note bench $? ;;
    bench4)
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29704 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu > $OUT/bench_n4.json 2> $OUT/bench_n4.err; note bench4 $? ;;
    scale)
      for k in 1 2 4 8; do
        [ $k -le $N ] || continue
        if [ $k -eq 1 ]; then timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu --no-sweep > $OUT/scale_n$k.json 2> $OUT/scale_n$k.err
        else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $k --master-addr 127.0.0.1 --master-port 2980$k bench.py --gpus $k --steps 20 --warmup 3 --no-cpu --no-sweep > $OUT/scale_n$k.json 2> $OUT/scale_n$k.err; fi
        note scale$k $?
      done ;;
    benchc4)
      timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29705 bench.py --gpus $N --workload c4 --steps 5 --warmup 3 --no-cpu --no-sweep > $OUT/bench_c4_n$N.json 2> $OUT/bench_c4_n$N.err; note benchc4 $? ;;
    multi)
      timeout 600 python tools/multi_bench.py --gpus $N > $OUT/multi_n$N.json 2> $OUT/multi_n$N.err; note multi $? ;;
  esac
done
cat $OUT/progress.txt
[ -f $OUT/pytest_two.txt ] && tail -15 $OUT/pytest_two.txt
for f in $OUT/*.json; do echo "== $f"; cut -c1-1200 $f; done
for f in $OUT/*.err; do [ -s $f ] && { echo "== $f"; tail -5 $f; }; done
exit 0

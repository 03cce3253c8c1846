#!/bin/bash
# One gpurun call: A/B of the experiment builds in build_exp/ against build_exp/base.so, then (with the winner installed as the
# in-tree library) the bench line, the GPU test suite, the ncu launch list and one full capture of span_kernel.
# Everything lands in gpurun_out/$TAG/.  Usage on the box: bash tools/gpu_round.sh <tag> <baseline> <variant> [<variant> ...]
set +e
TAG=${1:-r01_v25}; BASE=${2:-base}; shift 2
OUT=gpurun_out/$TAG
mkdir -p $OUT
date +%s > $OUT/t0
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,temperature.gpu --format=csv > $OUT/smi.txt 2>&1

# 1. A/B (one process per library; fingerprints of the full-size tables must agree with base.so)
for v in $([ -n "$SKIP_AB" ] || echo $BASE "$@"); do
  UPB_LIB=$PWD/build_exp/$v.so timeout 300 python tools/ab.py c3 ctl c2 > $OUT/ab_$v.jsonl 2> $OUT/ab_$v.err
  echo "ab $v rc=$? t=$(( $(date +%s) - $(cat $OUT/t0) ))" >> $OUT/progress.txt
done
[ -n "$SKIP_AB" ] || python - "$OUT" $BASE "$@" > $OUT/ab_verdict.json <<'PY'
import json, sys, os, shutil
out, base_name, variants = sys.argv[1], sys.argv[2], sys.argv[3:]
def rd(v):
    try:
        return {r["workload"]: r for r in map(json.loads, open(os.path.join(out, "ab_%s.jsonl" % v)))}
    except Exception:
        return {}
base = rd(base_name)
best, best_ms = base_name, base.get("c3 bw=50", {}).get("span_kernel_ms", 1e9)
verdict = {base_name: {k: (r["span_kernel_ms"], r["call_ms"], r["sha"]) for k, r in base.items()}}
for v in variants:
    r = rd(v)
    same = bool(base) and set(r) == set(base) and all(r[k]["sha"] == base[k]["sha"] for k in base)
    verdict[v] = {"identical_tables": same, **{k: (x["span_kernel_ms"], x["call_ms"], x["sha"]) for k, x in r.items()}}
    if same and r["c3 bw=50"]["span_kernel_ms"] < 0.99 * best_ms:
        best, best_ms = v, r["c3 bw=50"]["span_kernel_ms"]
verdict["winner"] = best
shutil.copyfile("build_exp/%s.so" % best, "unipeak_b200/libunipeak_b200.so")
print(json.dumps(verdict, indent=1))
PY
echo "verdict t=$(( $(date +%s) - $(cat $OUT/t0) ))" >> $OUT/progress.txt

# 2. bench line of the winner (both arms)
timeout 300 python bench.py > $OUT/bench_c3.json 2> $OUT/bench_c3.err
echo "bench rc=$? t=$(( $(date +%s) - $(cat $OUT/t0) ))" >> $OUT/progress.txt

# 3. the GPU test suite on the winner
timeout 600 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.txt 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - $(cat $OUT/t0) ))" >> $OUT/progress.txt

# 4. ncu: launch list of the bench command, then one full capture of span_kernel
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:upb:: --csv --log-file $OUT/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu > $OUT/ncu_list.log 2>&1
echo "ncu list rc=$? t=$(( $(date +%s) - $(cat $OUT/t0) ))" >> $OUT/progress.txt
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:upb::span_kernel -c 1 -f -o $OUT/span \
  python bench.py --steps 2 --warmup 1 --no-cpu > $OUT/ncu_full.log 2>&1
echo "ncu full rc=$? t=$(( $(date +%s) - $(cat $OUT/t0) ))" >> $OUT/progress.txt
timeout 200 python bench.py --impl reference --steps 1 --warmup 0 > $OUT/reference_arm.json 2> $OUT/reference_arm.err
echo "ref arm rc=$? t=$(( $(date +%s) - $(cat $OUT/t0) ))" >> $OUT/progress.txt
tail -3 $OUT/pytest_gpu.txt; cat $OUT/ab_verdict.json | head -60; cat $OUT/bench_c3.json | cut -c1-600; cat $OUT/progress.txt

#!/bin/bash
# Short gpurun call: A/B of the builds in build_exp/ only (tools/ab.py per library), results in gpurun_out/$TAG/.
# usage on the box: bash tools/gpu_ab.sh <tag> "<ab.py workloads>" <variant> [<variant> ...]
set +e
TAG=$1; WL=$2; shift 2
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in "$@"; do
  UPB_LIB=$PWD/build_exp/$v.so timeout 300 python tools/ab.py $WL > $OUT/ab_$v.jsonl 2> $OUT/ab_$v.err || tail -3 $OUT/ab_$v.err
done
python - $OUT "$@" <<'PY'
import json, sys, os
out, vs = sys.argv[1], sys.argv[2:]
tab = {v: {r["workload"]: r for r in map(json.loads, open(os.path.join(out, "ab_%s.jsonl" % v)))} for v in vs}
for w in tab[vs[0]]:
    print("%-22s" % w, "  ".join("%s %.3f/%.3f %s" % (v, tab[v].get(w, {}).get("span_kernel_ms", -1), tab[v].get(w, {}).get("call_ms", -1), tab[v].get(w, {}).get("sha", "-")[:6]) for v in vs))
PY

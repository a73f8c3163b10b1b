#!/bin/bash
# bench.py over the BASELINE configs (workload x level), one JSON line each -> gpurun_out/matrix_<tag>.jsonl
#   bash tools/bench_matrix.sh <tag> [extra bench.py args]
tag=$1; shift
out=gpurun_out/matrix_${tag}.jsonl; : > $out
for wl in text random zeros pattern; do for lv in 2 1 0; do
  if [ $lv != 2 ] && [ $wl != text ] && [ $wl != random ]; then continue; fi
  echo "== $wl L$lv" >&2
  timeout 600 python bench.py --workload $wl --level $lv --steps 5 --warmup 3 "$@" 2>>gpurun_out/matrix_${tag}.err | tail -1 >> $out
done; done
python - <<PY
import json
for l in open("$out"):
    d = json.loads(l)
    print(d["config"]["workload"][:34], "L", d["config"]["level"], "dev", d["value"], "e2e", (d.get("e2e") or {}).get("value"), d.get("stage_ms_per_step"), "ok" if (d.get("verified") or {}).get("all_ranks_ok") else d.get("verified"))
PY

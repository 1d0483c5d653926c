#!/bin/bash
# A/B timing inside ONE gpurun call (boxes differ by several %): alternates the variants given as "name:ENV=..;flags" arguments.
# usage: tools/ab_bench.sh rounds "tiles:;" "ring:;--ring" "plain:;--plain"
rounds=$1; shift
for r in $(seq 1 $rounds); do
  for v in "$@"; do
    name=${v%%:*}; rest=${v#*:}; envs=${rest%%;*}; flags=${rest#*;}
    env $envs python bench.py --steps 4 --warmup 3 --no-cpu-baseline $flags > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_$name.json"))
    print("$name", "round", $r, "value", round(d["value"],1), "avg_launch_ms", round(d["roofline"]["avg_launch_ms"],3), "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name", "FAILED", e)
PY
  done
done

#!/bin/bash
# GPU box: A/B of the default bench line with an environment switch on/off, alternating runs on one box.
# usage: ab_env.sh VAR [runs]      e.g. ab_env.sh HGRU_NO_CHAIN
var=$1; runs=${2:-2}
mkdir -p gpurun_out
for i in $(seq 1 $runs); do
  env $var=1 timeout 200 python bench.py --no-stages --no-cpu-baseline > gpurun_out/bench_ab_${var}_on_$i.json 2> gpurun_out/bench_ab_on.err
  timeout 200 python bench.py --no-stages --no-cpu-baseline > gpurun_out/bench_ab_${var}_off_$i.json 2> gpurun_out/bench_ab_off.err
done
python - "$var" "$runs" <<'PY'
import json, sys
var, runs = sys.argv[1], int(sys.argv[2])
for i in range(1, runs + 1):
    for st in ("on", "off"):
        try:
            d = json.load(open("gpurun_out/bench_ab_%s_%s_%d.json" % (var, st, i)))
            r = d["roofline"]
            print("%s=%s run %d: fps %.0f  ms/step %.3f  e2e %.0f  frac %.3f  avg_launch_ms %.4f" % (
                var, "1" if st == "on" else "-", i, d["value"], d["ms_per_step"], d["e2e"]["value"], r["frac"], r["avg_launch_ms"]))
        except Exception as e:
            print(var, st, i, "ERR", e)
PY
tail -n 3 gpurun_out/bench_ab_off.err

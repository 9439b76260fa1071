#!/bin/bash
# GPU box: launch chaining on/off A/B of the default bench line, alternating runs on one box.
mkdir -p gpurun_out
for i in 1 2; do
  HGRU_NO_CHAIN=1 timeout 200 python bench.py --no-stages --no-cpu-baseline > gpurun_out/bench_ab_nochain_$i.json 2> gpurun_out/bench_ab_nochain.err
  timeout 200 python bench.py --no-stages --no-cpu-baseline > gpurun_out/bench_ab_chain_$i.json 2> gpurun_out/bench_ab_chain.err
done
python - <<'PY'
import json
for f in ("nochain_1", "chain_1", "nochain_2", "chain_2"):
    try:
        d = json.load(open("gpurun_out/bench_ab_%s.json" % f))
        r = d["roofline"]
        print(f, "fps %.0f  ms/step %.3f  e2e %.0f  frac %.3f  avg_launch_ms %.4f  launches %d" % (
            d["value"], d["ms_per_step"], d["e2e"]["value"], r["frac"], r["avg_launch_ms"], r["launches_per_step"]))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -n 3 gpurun_out/bench_ab_chain.err

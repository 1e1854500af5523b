#!/bin/bash
# tuning sweep: canvases per warp in expand_kernel (SPIRAL_EXPAND_LANES), B=1024 and B=16384
for B in 1024 16384; do
  for L in 1 2 4 8 32; do
    SPIRAL_EXPAND_LANES=$L python bench.py --steps 2 --warmup 3 --no_cpu_baseline --canvases $B 2>/dev/null > /tmp/o.json
    python - <<PY
import json
d=json.load(open("/tmp/o.json"))
print("B=$B lanes=$L env-steps/s=%d ms_expand=%.3f ms_composite=%.3f" % (d["value"], d["roofline"]["k1_pair"]["ms_expand"], d["roofline"]["k1_pair"]["ms_composite"]))
PY
  done
done

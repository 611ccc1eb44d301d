#!/bin/bash
# A/B of the conv_tc A-tile modes inside one GPU session (same box, same clocks)
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_dcn.py tests/test_gpu_e2e.py -x -q 2>&1 | tail -3
for m in 1 0 1 0; do
  python bench.py --steps 2 --warmup 3 --pairs 16 --conv-mode $m --no-cpu-baseline --no-kernels 2>/dev/null > /tmp/ab_$m.json
  python - <<PY
import json
d = json.loads(open("/tmp/ab_$m.json").read().strip().splitlines()[-1])
print("mode $m", round(d["value"], 1), "pairs/s  conv_tc", round(d["roofline"]["achieved"], 1), "TF/s  share", round(d["roofline"]["share_of_step"], 3))
PY
done

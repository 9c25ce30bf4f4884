#!/bin/bash
# A/B commands for the first GPU call of the next round (each line prints conformers/s and the per-kernel ms of one evaluation).
# Usage on the GPU box:  bash tools/ab_next_round.sh
set -u
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 400 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - "$name" <<'PY'
import json, sys
d = json.load(open("gpurun_out/ab_%s.json" % sys.argv[1]))
print(sys.argv[1], round(d["value"], 2), "conformers/s", "fallbacks", d["range_fallbacks"], d["kernel_ms_per_forward"])
PY
}
mkdir -p gpurun_out
run default AGD_F16_WS=0
run ws AGD_F16_WS=1          # warp-specialised CFConv kernel (tc_filter16_ws.cu): functionally validated, never timed

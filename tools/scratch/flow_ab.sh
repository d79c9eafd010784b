#!/bin/bash
# A/B sweep of the FlowAlign kernels over the tuning builds in gpurun_ab/ (run on the GPU box)
for so in gpurun_ab/libvosd_*.so; do
  tag=$(basename $so .so)
  VOSD_B200_LIB=$so python tools/microbench.py --only flow --iters 15 --out gpurun_out/ab_$tag.json > /dev/null 2>&1
  python - <<PY
import json
r = {x["kernel"]: x["ms_cold_median"] for x in json.load(open("gpurun_out/ab_$tag.json"))["results"]}
print("%-16s fwd %.3f  fwd_cvt %.3f  fwd_fp32 %.3f | bwd %.3f  bwd_cvt %.3f   (4 frames, ms, L2 flushed)" % ("$tag", r["flowalign_fwd_5lvl_4frames"], r["flowalign_fwd_cvtvariant_5lvl_4frames"], r["flowalign_fwd_fp32variant_5lvl_4frames"], r["flowalign_bwd_5lvl_4frames"], r["flowalign_bwd_cvtvariant_5lvl_4frames"]))
PY
done

#!/bin/bash
# A/B sweep of the channels-last RoIAlign forward over the tuning builds in gpurun_ab/ (run on the GPU box)
for so in gpurun_ab/libvosd_*.so; do
  tag=$(basename $so .so)
  VOSD_B200_LIB=$so python tools/microbench.py --only nhwc --iters 12 --out gpurun_out/abn_$tag.json > /dev/null 2>&1
  VOSD_B200_LIB=$so python bench.py --no-cpu-baseline --steps 10 > gpurun_out/abn_bench_$tag.json 2>/dev/null
  python - <<PY
import json
r = {x["kernel"]: x["ms_cold_median"] for x in json.load(open("gpurun_out/abn_$tag.json"))["results"]}
b = json.load(open("gpurun_out/abn_bench_$tag.json"))["alt_layout"]
print("%-16s micro 7x7 %.3f (nchw %.3f)  14x14 %.3f | bench channels-last: box %.3f mask %.3f step %.3f ms" % ("$tag",
      r["bench_roialign_fwd_7x7_10000rois_channels_last_tma"], r["bench_roialign_fwd_7x7_10000rois_nchw"],
      r["bench_roialign_fwd_14x14_1000rois_channels_last_tma"], b["stages_ms"]["roialign_box"], b["stages_ms"]["roialign_mask"], b["ms_per_step"]))
PY
done

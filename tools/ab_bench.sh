#!/bin/bash
# On the GPU box: per-dataset LZ throughput with every library variant under build/variants/ (tools/build_variants.sh).
#   bash tools/ab_bench.sh <tag> [datasets] [codecs]
set -u
TAG=${1:-ab}
DS=${2:-tabular_f32:0,tabular_f32:1,snappy_synth,sorted_i64,tabular_f32,lz4_mixed}
CODECS=${3:-lz4,snappy}
O=gpurun_out/iter
mkdir -p $O
cp nvcomp_b200/lib/libnvcomp.so /tmp/libnvcomp_base.so
for d in build/variants/*/; do
  name=$(basename $d)
  cp $d/libnvcomp.so nvcomp_b200/lib/libnvcomp.so
  timeout 400 python tools/quick_bench.py --codecs $CODECS --datasets $DS > $O/${TAG}_$name.jsonl 2> $O/${TAG}_$name.err
  echo "== $name"; python - <<PY
import json
for l in open("$O/${TAG}_$name.jsonl"):
    d=json.loads(l); print(f"  {d['codec']:7s} {d['dataset']:16s} {d['decomp_GBps']:8.1f} ok={d['ok']}")
PY
done
cp /tmp/libnvcomp_base.so nvcomp_b200/lib/libnvcomp.so

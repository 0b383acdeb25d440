#!/bin/bash
# Kernel-iteration run on a GPU box: LZ + typed parity tests, per-dataset throughput of the LZ and Cascaded decoders,
# one full ncu capture.   bash tools/gpu_iter.sh <tag> [ncu_codec] [ncu_dataset] [ncu_kernel_regex]
set -u
TAG=${1:-it}
NC=${2:-cascaded}
ND=${3:-sorted_i64}
NK=${4:-${NC}_decompress}
O=gpurun_out/iter
mkdir -p $O
echo "== pytest lz + typed + fuzz"
timeout 1200 python -m pytest tests/test_lz_gpu.py tests/test_typed_gpu.py tests/test_fuzz_gpu.py -x -q -m gpu > $O/${TAG}_pytest.log 2>&1; tail -3 $O/${TAG}_pytest.log
echo "== per-dataset"
timeout 600 python tools/quick_bench.py --codecs lz4,snappy \
  --datasets runlength_i32,tabular_f32:0,tabular_f32:1,tabular_f32:2,tabular_f32:3,snappy_synth,sorted_i64,tabular_f32,lz4_mixed \
  > $O/${TAG}_lz_per_dataset.jsonl 2> $O/${TAG}_lz_per_dataset.err
cut -c1-200 $O/${TAG}_lz_per_dataset.jsonl
timeout 300 python tools/quick_bench.py --codecs cascaded,ans,bitcomp --datasets sorted_i64,runlength_i32,lowentropy_bytes \
  > $O/${TAG}_typed_per_dataset.jsonl 2> $O/${TAG}_typed_per_dataset.err
cut -c1-200 $O/${TAG}_typed_per_dataset.jsonl; tail -2 $O/${TAG}_typed_per_dataset.err
echo "== ncu"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:${NK} -c 1 -f \
  -o $O/${TAG}_${NC} python tools/quick_bench.py --codecs $NC --datasets $ND --iters 2 --no-verify > $O/${TAG}_ncu.log 2>&1
ls -la $O/${TAG}_${NC}.ncu-rep 2>&1 | cut -c1-120

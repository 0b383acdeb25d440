#!/bin/bash
# Collect the per-round evidence on a GPU box (run through gpurun from the repo root):
#   bash tools/collect_evidence.sh [tag]        -> gpurun_out/evidence/<tag>_*
# Then, back in the build container: python tools/publish_evidence.py [tag] copies / summarises into profiles/.
set -u
TAG=${1:-r2}
O=gpurun_out/evidence
mkdir -p $O
declare -A DS=( [snappy]=tabular_f32 [lz4]=lz4_mixed [cascaded]=sorted_i64 [bitcomp]=sorted_i64 [ans]=lowentropy_bytes )
CODECS="snappy lz4 cascaded bitcomp ans"

echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; tail -2 $O/${TAG}_pytest_gpu.log
echo "== smoke"
timeout 300 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log

echo "== bench (contract line per codec, reference arm)"
for c in $CODECS; do
  timeout 500 python bench.py --codec $c 2> $O/${TAG}_bench_$c.err | tail -1 > $O/${TAG}_bench_$c.json
  cut -c1-160 $O/${TAG}_bench_$c.json
done
timeout 400 python bench.py --impl reference 2> $O/${TAG}_bench_reference.err | tail -1 > $O/${TAG}_bench_reference.json
cut -c1-160 $O/${TAG}_bench_reference.json

echo "== per-dataset LZ numbers"
timeout 500 python tools/quick_bench.py --codecs lz4,snappy \
  --datasets runlength_i32,tabular_f32:0,tabular_f32:1,tabular_f32:2,tabular_f32:3,snappy_synth,sorted_i64 \
  > $O/${TAG}_lz_per_dataset.jsonl 2> $O/${TAG}_lz_per_dataset.err
wc -l $O/${TAG}_lz_per_dataset.jsonl

echo "== ncu: launch list of the default bench"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $O/${TAG}_bench_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > $O/${TAG}_launches.log 2>&1
wc -l $O/${TAG}_bench_launches.csv

echo "== ncu: one full capture per decode kernel"
declare -A KN=( [snappy]=snappy_decompress_v2 [lz4]=lz4_decompress_v2 [cascaded]=cascaded_decompress [bitcomp]=bitcomp_decompress [ans]=ans_decompress )
for c in $CODECS; do
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:${KN[$c]} -c 1 -f \
    -o $O/${TAG}_${c}_${DS[$c]} python tools/quick_bench.py --codecs $c --datasets ${DS[$c]} --iters 2 --no-verify \
    > $O/${TAG}_ncu_$c.log 2>&1
  ls -la $O/${TAG}_${c}_${DS[$c]}.ncu-rep 2>&1 | cut -c1-120
done
# DRAM traffic of the light (direct) kernel of the two LZ codecs on the bench workloads: the bench's `traffic` is the
# sum over both decode kernels of one DecompressAsync
for c in snappy lz4; do
  timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:${c}_decompress_light -c 1 --csv \
    --log-file $O/${TAG}_traffic_light_$c.csv python tools/quick_bench.py --codecs $c --datasets ${DS[$c]} --iters 2 --no-verify \
    > $O/${TAG}_ncu_light_$c.log 2>&1
  tail -2 $O/${TAG}_traffic_light_$c.csv | cut -c1-200
done
# the dense block decoder alone on the survey's cfg2-ii column, and the light (direct) kernel on run-length data
timeout 500 ncu --set full --clock-control none --import-source on -k regex:snappy_decompress_v2 -c 1 -f \
  -o $O/${TAG}_snappy_price_walk python tools/quick_bench.py --codecs snappy --datasets tabular_f32:0 --iters 2 --no-verify \
  > $O/${TAG}_ncu_snappy_pw.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:lz4_decompress_light -c 1 -f \
  -o $O/${TAG}_lz4_runlength_i32 python tools/quick_bench.py --codecs lz4 --datasets runlength_i32 --iters 2 --no-verify \
  > $O/${TAG}_ncu_lz4_rl.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:snappy_decompress_light -c 1 -f \
  -o $O/${TAG}_snappy_runlength_i32 python tools/quick_bench.py --codecs snappy --datasets runlength_i32 --iters 2 --no-verify \
  > $O/${TAG}_ncu_snappy_rl.log 2>&1
ls -la $O/${TAG}_snappy_price_walk.ncu-rep $O/${TAG}_lz4_runlength_i32.ncu-rep $O/${TAG}_snappy_runlength_i32.ncu-rep 2>&1 | cut -c1-120

echo "== compute-sanitizer memcheck"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_fuzz_gpu.py -x -q -m gpu \
  > $O/${TAG}_memcheck_fuzz.log 2>&1; echo "fuzz rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $O/${TAG}_memcheck_fuzz.log | tail -2
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_lz_gpu.py tests/test_typed_gpu.py -x -q -m gpu \
  > $O/${TAG}_memcheck_parity.log 2>&1; echo "parity rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $O/${TAG}_memcheck_parity.log | tail -2

set -u
O=gpurun_out/iter; mkdir -p $O
timeout 900 python -m pytest tests/test_typed_gpu.py tests/test_fuzz_gpu.py -x -q -m gpu > $O/c8_pytest.log 2>&1; tail -3 $O/c8_pytest.log
timeout 300 python tools/quick_bench.py --codecs cascaded --datasets sorted_i64,runlength_i32,tabular_f32 > $O/c8_casc.jsonl 2> $O/c8_casc.err; cut -c1-220 $O/c8_casc.jsonl; tail -2 $O/c8_casc.err
timeout 500 ncu --set full --clock-control none --import-source on -k regex:cascaded_decompress -c 1 -f \
  -o $O/c8_cascaded python tools/quick_bench.py --codecs cascaded --datasets sorted_i64 --iters 2 --no-verify > $O/c8_ncu.log 2>&1
ls -la $O/c8_cascaded.ncu-rep | cut -c1-100

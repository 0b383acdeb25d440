set -u
O=gpurun_out/iter; mkdir -p $O
timeout 600 python -m pytest tests/test_lz_gpu.py tests/test_fuzz_gpu.py tests/test_multi_device_gpu.py -x -q -m gpu > $O/c11_pytest.log 2>&1; tail -3 $O/c11_pytest.log
bash tools/ab_bench.sh c11 tabular_f32,lz4_mixed,runlength_i32,tabular_f32:0,tabular_f32:2 lz4,snappy

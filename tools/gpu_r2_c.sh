#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_spd_generator.py -m gpu -q -k "persistent or spd or generator or few_sms or cooperative" > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
timeout 600 python tools/small_n_bench.py 2048 1024 512 3000 4096 > gpurun_out/r2c_small_n.log 2>&1
timeout 900 python tools/spd_bench.py 2048 8192 16384 > gpurun_out/r2c_spd.log 2>&1
tail -4 gpurun_out/r2c_pytest.log; cat gpurun_out/r2c_small_n.log gpurun_out/r2c_spd.log

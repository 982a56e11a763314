#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
timeout 600 python tools/small_n_bench.py 2048 1024 512 256 3000 > gpurun_out/r2d_small_n.log 2>&1
LAMCG_SPD_VERBOSE=1 timeout 900 python tools/spd_bench.py 2048 8192 16384 > gpurun_out/r2d_spd.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 --no-reference-gpu --no-ncu-traffic > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
grep -n "passed\|failed\|FAILED\|rc=" gpurun_out/r2d_pytest.log | tail; cat gpurun_out/r2d_small_n.log gpurun_out/r2d_spd.log; tail -3 gpurun_out/r2d_bench.err

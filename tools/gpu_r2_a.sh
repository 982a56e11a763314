#!/bin/bash
# round-2 GPU session A: full GPU test suite, cluster microbenchmark, K1 sweep, ingest sweep, short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_env.txt; nproc >> gpurun_out/r2a_env.txt; free -g >> gpurun_out/r2a_env.txt; df -h /tmp >> gpurun_out/r2a_env.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 120 tools/cluster_exchange.out > gpurun_out/r2a_cluster_exchange.log 2>&1
timeout 600 python tools/gemv_sweep.py 100000 50000 10000 --variants 36,46,32,42,11,2 > gpurun_out/r2a_sweep_1gpu.log 2>&1
timeout 300 python tools/gemv_sweep.py 100000 --ranks 8 --variants 36,46,32,42 --reps 50 > gpurun_out/r2a_sweep_r8.log 2>&1
timeout 300 python tools/ingest_bench.py 20000 > gpurun_out/r2a_ingest.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -3 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_cluster_exchange.log

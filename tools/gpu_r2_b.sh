#!/bin/bash
# round-2 GPU session B: full GPU test suite (no -x), all-gather microbenchmark, the new bench.py
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
timeout 120 tools/allgather_bench.out > gpurun_out/r2b_allgather.log 2>&1
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?" >> gpurun_out/r2b_bench.err
timeout 900 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err
tail -5 gpurun_out/r2b_pytest.log; cat gpurun_out/r2b_allgather.log; tail -3 gpurun_out/r2b_bench.err

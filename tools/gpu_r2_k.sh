#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
timeout 600 python tools/small_n_bench.py 2048 1024 512 256 3000 3500 4096 > gpurun_out/r2k_small_n.log 2>&1
grep -n "passed\|failed\|FAILED\|rc=" gpurun_out/r2k_pytest.log | tail -5; grep "gen4 (gathered Ap) poll v4\|gen1 rows_smem=auto\|gen3" gpurun_out/r2k_small_n.log

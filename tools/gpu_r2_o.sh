#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
timeout 600 python tools/small_n_bench.py 256 512 1024 2048 3000 4096 5000 8192 10000 16384 > gpurun_out/r2o_small_n.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2o_smoke.log
grep -n "passed\|failed\|FAILED\|rc=" gpurun_out/r2o_pytest.log | tail -5; grep "persistent" gpurun_out/r2o_small_n.log; cat gpurun_out/r2o_smoke.log

#!/bin/bash
# round-2 last validation on 1 GPU with the final code: exactly the driver's commands
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2z_smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?" >> gpurun_out/r2z_bench.err
grep -n "passed\|failed\|FAILED\|rc=" gpurun_out/r2z_pytest.log | tail -4; cat gpurun_out/r2z_smoke.log | tail -3; tail -2 gpurun_out/r2z_bench.err

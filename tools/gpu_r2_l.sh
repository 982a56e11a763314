#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
timeout 600 python tools/small_n_bench.py 2048 1024 512 256 3000 4096 8192 > gpurun_out/r2l_small_n.log 2>&1
timeout 900 python bench.py > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?" >> gpurun_out/r2l_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2l_bench_ref.json 2> gpurun_out/r2l_bench_ref.err
python tools/persist_one.py 2048 300 4 > gpurun_out/r2l_persist_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:cg_persistent_v4 -c 1 -f -o gpurun_out/r02_persist_gen4_n2048 python tools/persist_one.py 2048 300 4 > gpurun_out/r2l_ncu_persist.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2l_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2l_smoke.log
grep -n "passed\|failed\|FAILED\|rc=" gpurun_out/r2l_pytest.log | tail -5; grep "gen4 (gathered Ap) poll v4" gpurun_out/r2l_small_n.log; tail -2 gpurun_out/r2l_bench.err; cat gpurun_out/r2l_smoke.log

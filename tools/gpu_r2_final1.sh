#!/bin/bash
# round-2 final validation on 1 GPU: the driver's own commands, then small-n numbers and the ncu capture of the gen-4 kernel
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2q_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2q_smoke.log
timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2q_bench_ref.json 2> gpurun_out/r2q_bench_ref.err
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?" >> gpurun_out/r2q_bench.err
timeout 600 python tools/small_n_bench.py 256 512 1024 2048 3000 4096 10000 > gpurun_out/r2q_small_n.log 2>&1
python tools/persist_one.py 2048 300 4 > gpurun_out/r2q_persist_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:cg_persistent_v4 -c 1 -f -o gpurun_out/r02_persist_gen4_n2048 python tools/persist_one.py 2048 300 4 > gpurun_out/r2q_ncu_persist.log 2>&1
grep -n "passed\|failed\|FAILED\|rc=" gpurun_out/r2q_pytest.log | tail -4; cat gpurun_out/r2q_smoke.log; tail -2 gpurun_out/r2q_bench.err; grep "persistent" gpurun_out/r2q_small_n.log

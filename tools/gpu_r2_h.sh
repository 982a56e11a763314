#!/bin/bash
# round-2 profiling session (1 GPU): un-profiled runs first, then ncu
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2h_bench_plain.json 2> gpurun_out/r2h_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_n100k.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2h_ncu_launches.log 2>&1
python tools/traffic_probe.py 100000 1 > gpurun_out/r2h_probe_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:lamcg_rowsweep_kernel -s 1 -c 1 -f -o gpurun_out/r02_k1_full_n100k python tools/traffic_probe.py 100000 1 > gpurun_out/r2h_ncu_k1.log 2>&1
python tools/persist_one.py 2048 300 4 > gpurun_out/r2h_persist_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:cg_persistent_v4 -c 1 -f -o gpurun_out/r02_persist_gen4_n2048 python tools/persist_one.py 2048 300 4 > gpurun_out/r2h_ncu_persist.log 2>&1
python tools/spd_one.py 8192 > gpurun_out/r2h_spd_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:gemm_f64_mma -s 40 -c 1 -f -o gpurun_out/r02_dmma_gemm_n8192 python tools/spd_one.py 8192 > gpurun_out/r2h_ncu_dmma.log 2>&1
# compute-sanitizer on the small-system paths (round 1: refused by the pool)
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "integer_inputs_bit_exact and (1185 or 4099 or 9-) or persistent_loop_generate_mode_vs_oracle and (147- or 1025-)" > gpurun_out/r02_sanitizer.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer.log
timeout 300 python tools/gen4_probe.py 2048 > gpurun_out/r2h_gen4_probe.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -5 gpurun_out/r02_sanitizer.log; grep "gen4 copies=2 poll=0" gpurun_out/r2h_gen4_probe.log

#!/bin/bash
# round 2: option matrix_f32 (fp32 matrix block under an fp64 solve) — its tests, the whole GPU suite again (K1 is templated on
# the storage type now), and the bench with the mixed-storage block
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mixed.py -q -m gpu > gpurun_out/r2m_pytest_mixed.log 2>&1; echo "mixed rc=$?" >> gpurun_out/r2m_pytest_mixed.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?" >> gpurun_out/r2m_bench.err
tail -15 gpurun_out/r2m_pytest_mixed.log; tail -4 gpurun_out/r2m_pytest.log; tail -2 gpurun_out/r2m_bench.err

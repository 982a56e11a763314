#!/bin/bash
# round-2 final 8-GPU session: multi-rank parity (2 ranks incl. the peer-timeout case, 8 ranks peer + NCCL), bench at N = 8 and N = 1 on the same box
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "ranked_solve and (peer-2 or peer-8 or nccl-8)" > gpurun_out/r2s_pytest_8gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_pytest_8gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2s_bench8.json 2> gpurun_out/r2s_bench8.err; echo "bench rc=$?" >> gpurun_out/r2s_bench8.err
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2s_bench1.json 2> gpurun_out/r2s_bench1.err; echo "bench rc=$?" >> gpurun_out/r2s_bench1.err
tail -n 3 gpurun_out/r2s_pytest_8gpu.log; tail -n 2 gpurun_out/r2s_bench8.err gpurun_out/r2s_bench1.err

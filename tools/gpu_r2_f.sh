#!/bin/bash
# 2-GPU session: multi-rank tests (peer + NCCL, fused K2+K3, resume/checkpoint, C++ drivers), bench at N=2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2f_topo.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_resume.py -m gpu -q > gpurun_out/r2f_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2f_bench2.json 2> gpurun_out/r2f_bench2.err; echo "bench rc=$?" >> gpurun_out/r2f_bench2.err
tail -5 gpurun_out/r2f_pytest_2gpu.log; tail -3 gpurun_out/r2f_bench2.err; head -c 1500 gpurun_out/r2f_bench2.json

#!/bin/bash
# 8-GPU session: multi-rank tests at 4 and 8 ranks, bench at N = 8, 4 (+1 for the same-box efficiency), SPD generator check on GPU 0
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2g_topo.txt 2>&1; free -g >> gpurun_out/r2g_topo.txt; nproc >> gpurun_out/r2g_topo.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "ranked_solve" > gpurun_out/r2g_pytest_8gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest_8gpu.log
for N in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2g_bench$N.json 2> gpurun_out/r2g_bench$N.err; echo "bench rc=$?" >> gpurun_out/r2g_bench$N.err
done
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-reference-gpu > gpurun_out/r2g_bench1.json 2> gpurun_out/r2g_bench1.err; echo "bench rc=$?" >> gpurun_out/r2g_bench1.err
LAMCG_COMM=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 5 --warmup 3 --no-extras > gpurun_out/r2g_bench8_nccl.json 2> gpurun_out/r2g_bench8_nccl.err
timeout 600 python -m pytest tests/test_gpu_spd_generator.py -m gpu -q > gpurun_out/r2g_pytest_spd.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest_spd.log
LAMCG_SPD_VERBOSE=1 timeout 600 python tools/spd_bench.py 2048 16384 > gpurun_out/r2g_spd.log 2>&1
tail -3 gpurun_out/r2g_pytest_8gpu.log gpurun_out/r2g_pytest_spd.log; tail -2 gpurun_out/r2g_bench*.err; cat gpurun_out/r2g_spd.log

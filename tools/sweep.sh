#!/bin/bash
# Result sweeps in the style of the reference's SLURM scripts (TESTS/GPU_SCRIPTS/*.sh, TESTS/CPU_SCRIPTS/CPU_1_NODE_gen.sh:25-33),
# for one box: every line written is the reference's CSV line
#   n,ranks,threads,io_or_gen_s,avg_gemv_s,avg_iter_s,iters,rel_err,total_s
# usage: tools/sweep.sh OUT.txt [gen|file DIR]    env: GPUS="1 2 4 8"  SIZES="10000 20000 ..."  ITERS=15
set -u
OUT=${1:-sweep_results.txt}
MODE=${2:-gen}
DIR=${3:-io}
HERE="$(cd "$(dirname "$0")/.." && pwd)"
EXE=${EXE:-"$HERE/2024-eumaster4hpc-student-challenge_b200/test/test_CG_MultiGPUS_CUDA_MPI.out"}  # 9 fields; ..._NCCL.out adds comm_init_s after io_s (10)
GPUS=${GPUS:-1}
ITERS=${ITERS:-15}
SIZES=${SIZES:-"10000 20000 30000 40000 50000"}
{
  echo "-----------------------------------------------------"
  echo "-----------------B200_${MODE}-------------------------"
  echo "-----------------------------------------------------"
} >> "$OUT"
for n in $SIZES; do
  for g in $GPUS; do
    if [ "$MODE" = gen ]; then
      LAMCG_NGPUS=$g "$EXE" -s "$n" -i "$ITERS" -e 1e-9 -o /tmp/lamcg_sweep_sol.bin >> "$OUT"
    else
      LAMCG_NGPUS=$g "$EXE" -A "$DIR/matrix$n.bin" -b "$DIR/rhs$n.bin" -i 10000 -e 1e-9 -o /tmp/lamcg_sweep_sol.bin >> "$OUT"
    fi
  done
done
rm -f /tmp/lamcg_sweep_sol.bin

"""CPU test of LAM/src/B200/RankWorld.hpp — the fork-based stand-in for `mpirun -n P` that the C++ drivers use (the reference's
distributed drivers get rank/size from MPI_Init, test_CG_MultiGPUS_CUDA_NCCL.cpp:205-209).  The header has no CUDA dependency, so
the bootstrap collectives (barrier, allgather, bcast), the all-ranks agreement used before every collective step (all_ok: one
failing rank makes ALL ranks leave together instead of spinning in a barrier — round-1 ADVICE) and the exit-code aggregation are
exercised here with plain g++."""
import os
import subprocess

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR_DIR = os.path.join(REPO, "2024-eumaster4hpc-student-challenge_b200", "LAM", "src", "B200")

PROGRAM = r'''
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "RankWorld.hpp"
int main(int argc, char **argv)
{
    const int failing = argc > 1 ? std::atoi(argv[1]) : -1;   // rank whose "step" fails (-1: none)
    const int exit_rank = argc > 2 ? std::atoi(argv[2]) : -1; // rank that finishes with exit code 7
    LAM::RankWorld w = LAM::RankWorld::launch(0);             // LAMCG_NGPUS ranks, forked before anything else
    const int P = w.size(), r = w.rank();
    int mine = 100 + r, all[LAM::RankWorld::kMaxRanks] = {0};
    w.allgather(&mine, sizeof mine, all);
    for (int i = 0; i < P; ++i)
        if (all[i] != 100 + i) return w.finalize(50);
    char msg[32] = {0};
    if (r == P - 1) std::strcpy(msg, "from the last rank");
    w.bcast(msg, sizeof msg, P - 1);
    if (std::strcmp(msg, "from the last rank") != 0) return w.finalize(51);
    const bool ok = w.all_ok(r != failing);                   // every rank learns of the failure
    if (ok != (failing < 0 || failing >= P)) return w.finalize(52);
    for (int i = 0; i < 100; ++i) w.barrier();                // sense reversal survives many rounds
    if (r == 0) std::printf("P=%d all_ok=%d\n", P, (int)ok);
    return w.finalize(r == exit_rank ? 7 : 0);
}
'''


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    d = tmp_path_factory.mktemp("rankworld")
    src, out = d / "rw.cpp", d / "rw.out"
    src.write_text(PROGRAM)
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", HDR_DIR, str(src), "-o", str(out), "-lpthread"], check=True)
    return str(out)


@pytest.mark.parametrize("P", [1, 2, 3, 8])
def test_bootstrap_collectives_and_agreement(exe, P):
    env = dict(os.environ, LAMCG_NGPUS=str(P))
    res = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=60)
    assert res.returncode == 0 and res.stdout.strip() == f"P={P} all_ok=1", (res.returncode, res.stdout, res.stderr)
    for failing in range(P):                                   # whichever rank fails, everybody sees all_ok == false and exits cleanly
        res = subprocess.run([exe, str(failing)], capture_output=True, text=True, env=env, timeout=60)
        assert res.returncode == 0 and res.stdout.strip() == f"P={P} all_ok=0", (failing, res.returncode, res.stdout, res.stderr)


def test_worst_exit_code_is_reported_by_rank_0(exe):
    env = dict(os.environ, LAMCG_NGPUS="4")
    for exit_rank in (0, 2, 3):
        res = subprocess.run([exe, "-1", str(exit_rank)], capture_output=True, text=True, env=env, timeout=60)
        assert res.returncode == 7, (exit_rank, res.returncode)

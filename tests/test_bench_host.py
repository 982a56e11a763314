"""CPU tests of bench.py's host logic (no GPU): parsing of the ncu traffic capture, the two-solve slope of the reference GPU class,
the reference arm's JSON line on a small system.  The GPU legs of bench.py are exercised on the B200 box by the driver."""
import importlib.util
import json
import os
import subprocess
import sys
import types

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(REPO, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


NCU_CSV = '''==PROF== Connected to process 123
probe: rows 100000 variant 36 ms per launch 10.9
"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"
"0","123","python","box","lamcg_rowsweep_kernel","1","7","(512, 1, 1)","(148, 1, 1)","0","10.0","Command line profiler metrics","dram__bytes_read.sum","Gbyte","80.01"
"0","123","python","box","lamcg_rowsweep_kernel","1","7","(512, 1, 1)","(148, 1, 1)","0","10.0","Command line profiler metrics","dram__bytes_write.sum","Mbyte","9.5"
"1","123","python","box","lamcg_rowsweep_kernel","1","7","(512, 1, 1)","(148, 1, 1)","0","10.0","Command line profiler metrics","dram__bytes_read.sum","byte","80,003,000,000"
"1","123","python","box","lamcg_rowsweep_kernel","1","7","(512, 1, 1)","(148, 1, 1)","0","10.0","Command line profiler metrics","dram__bytes_write.sum","Kbyte","1,500"
'''


def test_traffic_capture_is_parsed_per_launch_with_units(bench, monkeypatch, tmp_path):
    monkeypatch.setattr(bench, "REPO", str(tmp_path))       # profiles/gemv_traffic_n1.json goes to a scratch tree
    os.makedirs(tmp_path / "profiles")
    os.makedirs(tmp_path / "tools")
    fake_ncu = tmp_path / "ncu"
    fake_ncu.write_text("#!/bin/sh\ncat <<'EOF_'\n" + NCU_CSV + "EOF_\n")
    fake_ncu.chmod(0o755)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.environ["PATH"])
    traffic, src = bench.measure_k1_traffic(100000, 1, 0)
    assert traffic == pytest.approx(80_003_000_000 + 1_500_000)   # the LAST captured launch: bytes + kilobytes, thousands separators
    assert "measured in this run" in src
    saved = json.load(open(tmp_path / "profiles" / "gemv_traffic_n1.json"))
    assert saved["ranks"] == 1 and saved["launches"][0] == pytest.approx(80.01e9 + 9.5e6)


def test_traffic_capture_failure_is_reported_not_raised(bench, monkeypatch, tmp_path):
    fake_ncu = tmp_path / "ncu"
    fake_ncu.write_text("#!/bin/sh\necho '==ERROR== ERR_NVGPUCTRPERM' >&2\nexit 1\n")
    fake_ncu.chmod(0o755)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.environ["PATH"])
    traffic, src = bench.measure_k1_traffic(100000, 8, 0)
    assert traffic is None and "no metrics" in src


def test_reference_gpu_loop_time_is_the_slope_of_two_solves(bench, monkeypatch):
    """The reference class re-uploads A in every solve(): loop time = (t(K1) - t(K0)) / (K1 - K0) with the faster of two runs each."""
    calls = []

    def fake_ref_gpu_solve(variant, ks, eps, n=None, timeout=None, **kw):
        calls.append((variant, tuple(ks), n))
        per_it, setup = (0.0095, 2.5) if n == 50000 else (0.031, 7.5)
        jitter = [0.0, 0.4, 0.9, 0.1]                       # upload noise: the minimum per iteration count must be used
        return [{"max_iters": k, "seconds": setup + per_it * k + jitter[i], "iters": k + 1, "rel": 1e-5} for i, k in enumerate(ks)]

    fake_oracle = types.SimpleNamespace(ref_gpu_available=lambda v: True, ref_gpu_solve=fake_ref_gpu_solve)
    monkeypatch.setitem(sys.modules, "oracle", fake_oracle)
    monkeypatch.setattr(bench, "host_mem_gb", lambda: (200.0, 150.0))
    out = bench.reference_gpu_block({"configs[1] generate n=50000 -i 1000": {"ms_per_iteration": 2.76}}, 91.0)
    assert [c[2] for c in calls] == [50000, 100000] and all(len(c[1]) == 4 for c in calls)
    s50, s100 = out["systems"]
    # min over the two runs of each count: (2.5 + 220 * 0.0095 + 0.1) - (2.5 + 20 * 0.0095 + 0.0) over 200 iterations = 10.0 ms (a single pair would say 11.5)
    assert s50["reference_ms_per_iteration"] == pytest.approx(10.0, abs=1e-9)
    assert s50["loop_speedup"] == pytest.approx(s50["reference_ms_per_iteration"] / 2.76)
    assert s100["this_library_ms_per_iteration"] == pytest.approx(1e3 / 91.0)
    monkeypatch.setattr(bench, "host_mem_gb", lambda: (64.0, 50.0))                               # too little host memory for n = 100000
    calls.clear()
    out = bench.reference_gpu_block({}, 91.0)
    assert [c[2] for c in calls] == [50000] and out["systems"][0]["this_library_ms_per_iteration"] is None


def test_reference_arm_prints_the_contract_line_on_a_small_system():
    """`bench.py --impl reference` on the whole (small) system with the unmodified reference, when oracle/_ref is present."""
    import oracle
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built")
    res = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--n", "1500", "--steps", "2", "--warmup", "1",
                          "--ref-full-iters", "5"], capture_output=True, text=True, timeout=300, env=dict(os.environ, OMP_NUM_THREADS="2"))
    assert res.returncode == 0, res.stderr
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "iterations/s" and line["higher_is_better"] is True
    assert line["config"]["whole_system"] is True and line["cpu_baseline"]["same_system"] is True
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["x_rel_l2_vs_oracle"] <= 1e-12
    assert line["e2e"] == {"value": line["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}

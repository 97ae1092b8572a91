"""CPU: the reference arm of bench.py (the reference's own train step on the host cores) prints ONE JSON line on stdout with
the keys the driver reads; bench.py's module-level constants name the BASELINE metric."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="8")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--ref-budget", "12"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "sentences/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    assert "fwd+bwd" in d["metric"] and d["config"]["seq_len"] == 128 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    # "reference" = the unmodified reference from baseline/_ref (installed by baseline/install_ref.py), "port" = the oracle
    # restatement when that install is absent
    want = "reference" if os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "src", "models.py")) else "port"
    assert cb["kind"] == want and cb["cores"] >= 1 and cb["value"] == d["value"] and "train step" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "sentences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_committed_ncu_launch_list_matches_the_traffic_summary_bench_reads():
    """bench.py's roofline.traffic comes from profiles/r02_train_traffic.json; that file must be what tools/ncu_summary.py
    derives from the committed ncu launch list (303 GEMM-family launches of one train step, the halo conv included)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csv_path = os.path.join(root, "profiles", "r02_launches_train_B128_L128.ncu.csv")
    out = os.path.join(root, ".pytest_cache", "traffic_check.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run([sys.executable, os.path.join(root, "tools", "ncu_summary.py"), "launches", csv_path, out], check=True,
                   stdout=subprocess.DEVNULL)
    got = json.load(open(out))
    want = json.load(open(os.path.join(root, "profiles", "r02_train_traffic.json")))
    assert got["gemm_launches"] == want["gemm_launches"] == 303
    assert abs(got["gemm_dram_bytes_per_launch"] - want["gemm_dram_bytes_per_launch"]) < 1.0
    assert 0.5 < want["gemm_time_share"] < 0.75
    assert any("conv64_halo_kernel" in k for k in want["per_kernel"])

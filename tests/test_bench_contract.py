"""bench.py's reference arm runs on host cores only, so its JSON-line contract is testable without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    # OMP_NUM_THREADS=1 is what torchrun exports: the CPU arm must still use every core of the process
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--workload", "glow_mnist"], capture_output=True, text=True, timeout=600,
                         cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout                       # exactly one JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "inv-conv fwd+bwd images/s" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and d["gpu_launches"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["n_gpus"] == 1
    assert d["config"]["name"] == "glow_mnist" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["sample"]
    assert cb["cores"] == len(os.sched_getaffinity(0))
    assert "parallelism" in d and "parallelism" not in d["config"]       # the config is the same in both arms
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_default_workload_is_the_largest_single_gpu_configuration():
    import bench
    assert bench.WORKLOADS["glow_imagenet32"][0] == [(12, 16, 16, 3, 48), (24, 8, 8, 3, 48), (48, 4, 4, 3, 48)]
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert 'ap.add_argument("--workload", default="glow_imagenet32"' in src


def test_roofline_names_the_binding_term():
    import bench
    lat = {"ffma": 4.8, "ffma2": 4.8, "shfl": 24.0, "lds": 29.0, "sts_syncwarp_lds": 35.0, "sts_barsync8_lds": 58.6,
           "barsync8": 28.7, "fadd": 4.8}
    chain, what = bench.chain_cycles("wave<cg=12,k=3x3,cc=6,ns=4,vec=2,iters=1> slots=16", lat, 12, 3)
    assert 130 < chain < 170 and "reduce levels" in what
    r = bench.solve_roofline(9.0, 31, 2462784, 6.6e7, chain, what, 6548.5, 72.5, 1965.0)
    assert r["bound"] == "wavefront-latency" and abs(r["frac"] - r["terms_us"]["wavefront"] / 9.0) < 1e-12
    assert r["hbm"]["frac"] < 0.1 and r["fp32"]["frac"] < 0.2 and r["unit"] == "diagonals/us"
    big = bench.solve_roofline(14000.0, 127, 4.03e9, 1.04e12, chain, what, 6548.5, 72.5, 1965.0)
    assert big["bound"] == "fp32" and big["unit"] == "TFLOP/s"

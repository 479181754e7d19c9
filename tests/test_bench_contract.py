"""bench.py's JSON contract, as far as it can be checked without a GPU: the reference arm (the reference's CPU ExSUM on
the host cores) runs here, prints ONE JSON line with the keys the driver reads, honours --steps / --warmup, and carries
exactly the `config` our own arm prints (the driver compares the two arms)."""
import json
import os
import subprocess
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--log2n", "18", "--no-extras"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, p.stdout
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["steps"] == 2 and line["warmup"] == 1 and line["gpu_launches"] == 0
    assert line["unit"] == "GB/s" and line["value"] > 0 and line["dtype"] == "f64" and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the very config our arm prints for the same command line
    sys.path.insert(0, ROOT)
    import bench
    args = types.SimpleNamespace(op="exsum", dist="loguniform", log2n=18, fpe="3,4,8", early_exit=0)
    assert line["config"] == bench.config_of(args, 1)


def test_roofline_is_the_slower_of_two():
    sys.path.insert(0, ROOT)
    import bench
    n, peak, r = 1 << 30, 6551.4, 1.82e13
    hbm = bench.roofline_of("exsum", "loguniform", 8, False, n, 1.5, peak, "x", r, 7300.0)      # wide data: direct deposits, 4 DADD
    assert hbm["bound"] == "hbm" and hbm["fp64_instr_per_elem"] == 4 and abs(hbm["frac"] - (n * 8 / (peak * 1e9)) / 1.5e-3) < 1e-3
    fp = bench.roofline_of("exsum", "naive", 8, False, n, 3.0, peak, "x", r, 7300.0)            # 48 DADD per element
    assert fp["bound"] == "fp64" and fp["fp64_instr_per_elem"] == 48 and abs(fp["frac"] - (n * 48 / r) / 3.0e-3) < 1e-3
    assert bench.roofline_of("exsum", "naive", 3, False, n, 1.5, peak, "x", r, None)["bound"] == "hbm"      # 18 DADD: under the HBM time
    assert bench.roofline_of("exdot", "illcond", 8, False, n, 5.0, peak, "x", r, None)["fp64_instr_per_elem"] == 68

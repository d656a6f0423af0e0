"""CPU checks of the measurement helpers of bench.py (no GPU, no library call)."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_split_by_bound_classifies_launches_against_the_ridge(tmp_path):
    b = _bench()
    peaks = dict(tflops=1386.1, hbm=6536.4)
    # one compute-bound 3x3 convolution (K = 9 * 256), one HBM-bound 1x1 (K = 64) with BN-backward epilogue operands,
    # one weight gradient; columns: cls tag m n k launches total_ms tflops
    p = tmp_path / "shapes.txt"
    p.write_text("# header\n"
                 "0 259 102400 256 2304 10 1.0 0\n"      # tag 3 + CTA-pair bit 256
                 "0 142 819200 256 64 5 1.5 0\n"
                 "1 0 256 64 1638400 4 0.7 0\n")
    out = b.split_by_bound(str(p), peaks, 10.0)
    assert abs(out["ridge_flop_per_byte"] - 1386.1e12 / 6536.4e9) < 1e-6
    t, h = out["tensor"], out["hbm"]
    assert t["launches"] == 10 and h["launches"] == 9
    fl = 2.0 * 102400 * 256 * 2304 * 10
    assert abs(t["achieved"] - fl / 1.0e-3 / 1e12) < 1e-6 * t["achieved"]
    assert abs(t["share_of_step"] - 0.1) < 1e-12 and abs(h["share_of_step"] - 0.22) < 1e-12
    # HBM-bound bytes: fprop operands + output + three 2-byte epilogue operands; wgrad: dY and X once + fp32 dW
    by0 = 2.0 * (819200 * 64 + 256 * 64 + 819200 * 256) + 3 * 2.0 * 819200 * 256
    by1 = 2.0 * 1638400 * (256 + 64) + 4.0 * 256 * 64
    assert abs(h["achieved"] - (by0 * 5 + by1 * 4) / 2.2e-3 / 1e9) < 1e-6 * h["achieved"]
    assert 0 < t["frac"] < 2 and 0 < h["frac"] < 2


def test_measured_peaks_fallback_and_file():
    b = _bench()
    p = b.measured_peaks()
    assert p["tflops"] > 100 and p["hbm"] > 1000 and isinstance(p["source"], str)


def _run_watchdog_script(body):
    import subprocess, sys, textwrap
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = "import sys, time\nsys.path.insert(0, %r)\nimport bench\n" % root + textwrap.dedent(body)
    return subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=120)


def test_watchdog_prints_the_measured_line_when_a_later_leg_never_returns():
    """bench.py's contract: once `value` exists the JSON line comes out, whatever the optional legs after it do."""
    import json
    r = _run_watchdog_script("""
        wd = bench.Watchdog()
        wd.arm(dict(metric="m", value=1.5, unit="knees/s"), 1)
        wd.update(e2e=dict(value=1.0))
        time.sleep(60)          # a leg that never returns
        print("not reached")
    """)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1 and "not reached" not in r.stdout
    line = json.loads(lines[0])
    assert line["value"] == 1.5 and line["e2e"] == {"value": 1.0} and "watchdog" in line
    assert "most recent call first" in r.stderr      # the post-mortem names where every thread stood


def test_watchdog_stall_detector_and_disarm():
    import json
    r = _run_watchdog_script("""
        wd = bench.Watchdog()
        wd.arm(dict(metric="m", value=2.0), 3600)
        wd.stall_seconds = 1
        for _ in range(3):      # steps that finish keep it quiet
            wd.beat(); time.sleep(0.4)
        time.sleep(30)          # a step that does not finish
    """)
    assert r.returncode == 0 and json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])["value"] == 2.0
    r = _run_watchdog_script("""
        wd = bench.Watchdog()
        wd.arm(dict(metric="m", value=2.0), 1)
        wd.disarm()
        time.sleep(2.5)
        print("finished normally")
    """)
    assert r.returncode == 0 and "finished normally" in r.stdout and "{" not in r.stdout

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

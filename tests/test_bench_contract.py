"""CPU: the parts of bench.py that run without a GPU - the JSON-line plumbing, the clock sampler's degradation when
neither NVML nor nvidia-smi is there, the FLOP constants behind `roofline.achieved`, and the reference arm end to end on
a tiny sample (the contract asks for ONE JSON line on stdout whatever libraries print)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _bench():
    sys.path.insert(0, ROOT)
    import bench
    return bench


def test_flop_constants_match_the_survey():
    b = _bench()
    # 1,983,168 MAC per LR pixel = 247,896 FLOP per output pixel for EDSR x4
    assert b.EDSR_FLOP_PER_LR_PX == 2 * 1_983_168 and b.EDSR_FLOP_PER_LR_PX / 16 == 247_896
    head = 2 * 27 * 64
    body = 2 * 9 * 64 * 64 * (2 * 16 + 1)
    up = 2 * 9 * 64 * 256 * (1 + 4)
    tail = 2 * 9 * 64 * 3 * 16
    assert head + body + up + tail == b.EDSR_FLOP_PER_LR_PX and b.EDSR_HEAD_FLOP_PER_LR_PX == head


def test_clock_sampler_degrades_without_tools(monkeypatch):
    b = _bench()
    monkeypatch.setenv("PATH", "/nonexistent")
    monkeypatch.setitem(sys.modules, "pynvml", None)          # import pynvml -> ImportError
    with b.ClockSampler(0) as s:
        pass
    assert s.result["sm_mhz"] is None and s.result["reasons"] == []


def test_peaks_come_from_the_driver_file_or_the_fallback():
    b = _bench()
    p = b.measured_peaks()
    assert p["source"] in ("measured", "fallback") and p["tflops"] <= p["tflops_burst"] and p["hbm_gbs"] > 1000


def test_reference_arm_prints_exactly_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--batch", "1", "--tile", "24"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sr_output_megapixels_per_sec" and d["unit"] == "MP/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0 and d["value"] > 0

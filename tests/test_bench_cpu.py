"""bench.py host logic that does not need a GPU: the reference-arm JSON line (schema the driver parses), the
reference-equivalent FLOP counts, the clock sampler's fallbacks."""
import argparse
import json

import pytest

import bench


def _ref_line(monkeypatch, capsys, workload):
    calls = []

    def fake_factory(vol=bench.VOL, batch=1, threads=None, workload="z1200"):
        def step(v=vol):
            calls.append((tuple(v), batch, workload))
        return step

    monkeypatch.setattr(bench, "cpu_reference_step_factory", fake_factory)
    monkeypatch.delenv("RANK", raising=False)
    args = argparse.Namespace(gpus=1, steps=2, warmup=1, workload=workload)
    assert bench.run_reference_arm(args) == 0
    out = capsys.readouterr().out.strip().splitlines()
    assert len(out) == 1                                   # ONE JSON line
    return json.loads(out[0]), calls


@pytest.mark.parametrize("workload", ["z1200", "fc600"])
def test_reference_arm_line_schema(monkeypatch, capsys, workload):
    line, calls = _ref_line(monkeypatch, capsys, workload)
    assert line["impl"] == "reference" and line["unit"] == "volumes/s" and line["higher_is_better"] is True
    assert line["steps"] == 2 and line["warmup"] == 1 and line["n_gpus"] == 1 and line["vs_baseline"] is None
    assert line["dtype"] == "f32" and line["data"] == "synthetic" and line["gpu_launches"] == 0
    assert set(line["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]
    assert ("FC-latent" in line["metric"]) == (workload == "fc600")
    # one warm-up on a reduced volume, then exactly `steps` timed iterations at the full size
    assert len(calls) == 3 and calls[0][0] != bench.VOL and all(c[0] == bench.VOL for c in calls[1:])


def test_reference_arm_only_rank0_works(monkeypatch, capsys):
    monkeypatch.setenv("RANK", "1")
    monkeypatch.setattr(bench, "cpu_reference_step_factory", lambda **k: pytest.fail("rank 1 must not run the CPU arm"))
    assert bench.run_reference_arm(argparse.Namespace(gpus=2, steps=1, warmup=1, workload="z1200")) == 0
    assert capsys.readouterr().out == ""


def test_flop_counts():
    # headline: 32 encoder-or-decoder-sized passes of 227.20 GFLOP
    assert bench.GFLOP_PER_VOLUME_STEP == pytest.approx(32 * 227.2, rel=1e-3)
    g = bench.fc_gflop_per_volume_step((32, 64, 128, 256), 600, (5, 6, 5))
    assert 1800 < g < 2100
    # doubling every channel count quadruples the convolution FLOPs (the 1-channel stem / tail and the heads scale by 2)
    g2 = bench.fc_gflop_per_volume_step((64, 128, 256, 512), 600, (5, 6, 5))
    assert 3.8 < g2 / g < 4.0


def test_clock_sampler_without_nvidia_smi(monkeypatch):
    s = bench.ClockSampler(0)
    monkeypatch.setattr(bench.subprocess, "Popen", lambda *a, **k: (_ for _ in ()).throw(FileNotFoundError()))
    s.start()
    s.mark()
    out = s.stop()
    assert out["sm_mhz"] is None and out["reasons"] == ["nvidia-smi unavailable"]


def test_clock_sampler_window():
    s = bench.ClockSampler(0)

    class P:
        def terminate(self):
            pass

    s.proc = P()
    idle = "345, 1965, 140.0, Not Active, Not Active, Not Active, Not Active"
    busy = "1905, 1965, 950.0, Not Active, Not Active, Not Active, Active"
    s.lines = [idle] * 5
    s.mark()
    s.lines += [busy] * 3
    out = s.stop()
    assert out["sm_mhz"] == 1905 and out["samples"] == 3 and out["reasons"] == ["sw_power_cap"]
    assert out["window"] == "timed region"
    # a timed region shorter than two sampling periods falls back to everything sampled, and says so
    s2 = bench.ClockSampler(0)
    s2.proc = P()
    s2.lines = [busy] * 4
    s2.mark()
    s2.lines += [busy]
    out2 = s2.stop()
    assert out2["samples"] == 5 and out2["window"].startswith("warm-up")


def test_top_kernel_profile_is_read_from_profiles(tmp_path, monkeypatch):
    """roofline.traffic is not a constant in bench.py: it comes from the committed ncu summary of the current kernel
    (profiles/top_kernel_ncu.json), and is null when no capture is committed."""
    monkeypatch.setattr(bench, "TOP_KERNEL_PROFILE", str(tmp_path / "absent.json"))
    traffic, note = bench._top_kernel_profile()
    assert traffic is None and "absent" in note
    p = tmp_path / "top_kernel_ncu.json"
    p.write_text(json.dumps({"kernel": "conv3_kd3_kernel", "shape": "64->64 @ 8x80x96x80", "duration_us": 800.0,
                             "tensor_pipe_active_pct": 70.0, "dram_bytes_per_launch": 1.23e9, "commit": "abc1234",
                             "summary": "r02_ncu_top_kernel.md"}))
    monkeypatch.setattr(bench, "TOP_KERNEL_PROFILE", str(p))
    traffic, note = bench._top_kernel_profile()
    assert traffic == 1.23e9 and "abc1234" in note and "r02_ncu_top_kernel.md" in note


def test_hw_flop_factors():
    # Upsample(2) folded into the convolution executes 8 of the 27 reference taps per output voxel
    assert bench.HW_FLOP_FACTOR["upconv3_fprop"] == pytest.approx(8 / 27)
    assert bench.HW_FLOP_FACTOR["conv3_igemm"] == 1.0

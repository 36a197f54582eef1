"""bench.py's host logic, no GPU: the workload both arms report, the byte counts behind the roofline, the reference arm's
JSON line on a small problem (the oracle is the thing timed there -- the one place outside tests/ that may run it), and the rule
that a committed ncu traffic figure is only quoted for the kernel build it was captured on."""
import json
import os
import subprocess
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def args(**over):
    a = dict(n=512, smooth=2, box=32, scaling="weak", keep_b=False, smoother=1, fused_cfg=None, cpu_n=0, gpus=1, steps=1, warmup=0)
    a.update(over)
    return types.SimpleNamespace(**a)


def test_weak_scaling_domains_follow_survey_8d():
    """C5: 512^3 -> 512x512x1024 -> 512x1024x1024 -> 1024^3; C3 (strong): the same n^3 at every GPU count"""
    dims = {w: bench.workload_config(args(), w)["global_N"] for w in (1, 2, 4, 8)}
    assert dims == {1: [512, 512, 512], 2: [512, 512, 1024], 4: [512, 1024, 1024], 8: [1024, 1024, 1024]}
    assert all(bench.workload_config(args(scaling="strong"), w)["global_N"] == [512, 512, 512] for w in (1, 2, 4, 8))
    assert bench.workload_config(args(), 8)["global_cells"] == 8 * 512 ** 3


def test_both_arms_report_the_same_workload():
    """config is a pure function of the command line and the GPU count (the driver compares the two arms' config)"""
    a = args(gpus=4)
    c1, c2 = bench.workload_config(a, 4), bench.workload_config(a, 4)
    assert c1 == c2 and c1["n"] == 512 and c1["max_grid_size"] == 32 and "model" not in c1
    assert "512^3 cells per GPU" in c1["workload"] and "V(2,2)" in c1["workload"]
    assert bench.workload_config(args(keep_b=True), 1)["bCoef"] == "streamed"


def test_algorithmic_bytes():
    """48 B/cell per sweep with bCoef streamed, 40 with the all-ones bCoef dropped; a level of a V(2,2) cycle =
    four sweeps + restriction (33 / 25) + the zero fill, the prolongation's read and its read-modify-write (1 + 1 + 17)"""
    assert bench.algorithmic_bytes_per_cell(2, True) == (48, 48 * 4 + 52)
    assert bench.algorithmic_bytes_per_cell(2, False) == (40, 40 * 4 + 44)
    assert bench.algorithmic_bytes_per_cell(4, False)[1] == 40 * 8 + 44


def test_traffic_is_only_quoted_for_the_profiled_build(monkeypatch):
    p = os.path.join(ROOT, "profiles", bench.TRAFFIC_SOURCE)
    assert os.path.exists(p), "the ncu summary the bench quotes is committed"
    d = json.load(open(p))
    assert len(d["kernel_source_sha16"]) == 16 and len(bench.kernel_fingerprint()) == 16
    monkeypatch.setattr(bench, "kernel_fingerprint", lambda: d["kernel_source_sha16"])
    t = bench.profiled_traffic(args())
    assert t is not None and 5.0e9 < t < 6.0e9                                # ~5.36 GB per finest-level launch vs 5.13 GB algorithmic
    assert bench.profiled_traffic(args(n=256)) is None                        # another workload
    assert bench.profiled_traffic(args(keep_b=True)) is None
    monkeypatch.setattr(bench, "kernel_fingerprint", lambda: "0" * 16)         # the kernel source changed since the capture
    assert bench.profiled_traffic(args()) is None


def test_cpu_sample_shrinks_only_when_the_host_is_small(monkeypatch):
    monkeypatch.setattr(bench, "host_mem_available_gib", lambda: 256.0)
    assert bench.cpu_sample_n(args()) == 512
    monkeypatch.setattr(bench, "host_mem_available_gib", lambda: 16.0)
    assert bench.cpu_sample_n(args()) == 256
    monkeypatch.setattr(bench, "host_mem_available_gib", lambda: None)
    assert bench.cpu_sample_n(args(cpu_n=128)) == 128


def test_reference_arm_prints_one_json_line():
    """--impl reference on a small problem: ONE line on stdout with the contract's keys, every host core in use whatever
    OMP_NUM_THREADS says (torchrun exports 1), nothing of the GPU arm in it"""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "32", "--box", "16", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == "GDOF/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["gpu_launches"] == 0 and d["vs_baseline"] is None
    assert d["e2e"] == {"value": d["value"], "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["n_timed"] == 32
    assert cb["cores"] == len(os.sched_getaffinity(0))
    assert d["config"]["n"] == 32 and d["config"]["global_N"] == [32, 32, 32]

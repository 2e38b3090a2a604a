"""bench.py contract on the CPU: the reference arm (the oracle port timed on the host cores) prints ONE JSON line with the
same metric / unit / config keys as the CUDA arm, its own cpu_baseline and a zero-copy e2e object; ranks other than 0
print nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ, **(env_extra or {}))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.strip()]


def test_reference_arm_prints_one_json_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("ASD steps/sec") and d["unit"] == "steps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    # every step really runs one bounded sample: ms_per_step is ITS measured time (steps x ms_per_step = wall clock of the
    # loop), `value` is the full-step rate extrapolated by the stated factors and labelled as such
    assert d["value"] > 0 and d["extrapolated"] is True
    assert d["config"]["workload"].startswith("C2")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "render" in cb["sample"]
    assert cb["extrapolated"] is True and cb["full_step_seconds"] > cb["sample_seconds"] > 0
    assert abs(cb["full_step_seconds"] - 1.0 / d["value"]) < 1e-6 * cb["full_step_seconds"]
    assert 0.5 * cb["sample_seconds"] * 1e3 < d["ms_per_step"] < 3.0 * cb["sample_seconds"] * 1e3 + 5e3
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_never_maps_the_product_library():
    """The CPU baseline is the oracle alone: the process that runs it must not load libsdb200.so."""
    code = ("import sys, os; sys.argv=['bench.py','--impl','reference','--steps','1','--warmup','0']; "
            "sys.path.insert(0, %r); import bench; bench.main(); "
            "maps=open('/proc/self/maps').read(); assert 'libsdb200' not in maps, 'product library mapped'; "
            "assert 'scaledreamer_b200.lib' not in sys.modules or sys.modules['scaledreamer_b200.lib']._lib is None") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]


def test_reference_arm_is_silent_on_other_ranks():
    assert _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}) == []


def test_roofline_traffic_is_keyed_to_the_sources_it_was_captured_from(tmp_path, monkeypatch):
    """`roofline.traffic` comes from the committed ncu capture only for kernel groups whose CUDA sources still hash to what
    the capture recorded; a group whose sources changed reads as absent (null in the JSON line), never as a stale number."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import glob

    import bench
    import summarize_profiles as sp

    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")), key=os.path.getmtime)
    assert files, "no committed traffic capture"
    t = json.load(open(files[-1]))
    assert set(t["_group_sources_sha"]) == set(sp.SOURCE_GROUPS)
    now = bench.ncu_traffic()
    for group, keys in (("tensor", ["tensor"]), ("render", ["render_fwd", "render_bwd"]),
                        ("hyper_field", ["hyper_field_fwd", "hyper_field_bwd"])):
        same = t["_group_sources_sha"][group] == sp.lib_sources_sha(group)
        for k in keys:
            assert (k in now) == same, (group, k, same)
    # a group whose hash no longer matches disappears, the others stay
    real = sp.lib_sources_sha
    monkeypatch.setattr(sp, "lib_sources_sha", lambda g=None: "0" * 16 if g == "render" else real(g))
    changed = bench.ncu_traffic()
    assert "render_fwd" not in changed and "render_bwd" not in changed
    assert ("tensor" in changed) == ("tensor" in now)

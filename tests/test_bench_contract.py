"""bench.py's reference arm runs on host cores only, so its contract can be checked here: one JSON line with the keys the
driver reads, rank 0 alone printing it.  (The GPU arm's line is produced on the B200 box: profiles/*_bench.json.)"""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3",
                           "--ref-frames", "1"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mpix/s warped" and d["unit"] == "Mpix/s"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 1 and d["warmup"] >= 3 and d["n_gpus"] == 1
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and "1920x1080" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_headline_frames_have_distinct_points_and_one_window():
    """bench.py's per-frame destiny points (a different transform per frame, as in test/benchmark.js:96-113) all produce the
    1728 x 1080 window of BASELINE config 2, frame 0 being exactly config 2 — checked with the oracle's limits."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    import homography_js_b200 as hg
    from oracle import oracle as O
    O.build()
    wl = hg.workloads.projective_1080p()
    pts = bench.headline_points(wl, 130, 0)
    assert np.array_equal(pts[0], np.asarray(wl["dst"], np.float64))
    assert len({tuple(p) for p in pts[:64]}) == 64 and np.array_equal(pts[64], pts[0])   # period 64
    for f in (0, 1, 17, 63, 129):
        lim = O.transform_limits(O.projective_from_squares(wl["src"], pts[f]), wl["W"], wl["H"])
        assert [int(v) for v in lim] == [wl["x_off"], wl["y_off"], wl["o_w"], wl["o_h"]], (f, lim)
    # both arms print the same config dictionary (the driver compares them)
    assert bench.config_dict(wl) == bench.config_dict(hg.workloads.projective_1080p())

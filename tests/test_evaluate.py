"""CPU: accuracy-evaluation helpers (F4) -- known answers for ATE / RPE / compareGT."""
import importlib.util
import os

import numpy as np

from conftest import ROOT, load_golden


def _load():
    spec = importlib.util.spec_from_file_location("slam_b200_evaluate", os.path.join(ROOT, "slam-2d-lidar-scan_b200", "evaluate.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_ate_is_invariant_to_a_rigid_motion_and_rpe_sees_drift():
    E = _load()
    t = np.linspace(0, 4, 200)
    gt = np.stack([3 * np.cos(t), 3 * np.sin(t), t + np.pi / 2], 1)
    c, s = np.cos(0.7), np.sin(0.7)
    moved = gt.copy()
    moved[:, :2] = gt[:, :2] @ np.array([[c, -s], [s, c]]).T + [5.0, -2.0]
    moved[:, 2] += 0.7
    a = E.evaluate_trajectory(moved, gt)
    assert a["ate"]["rmse"] < 1e-12 and a["rpe"]["trans_rmse"] < 1e-12 and a["rpe"]["rot_rmse"] < 1e-12
    noisy = gt.copy()
    noisy[:, 0] += 0.1                                   # pure offset: removed by the alignment
    assert E.absolute_trajectory_error(noisy, gt)["rmse"] < 1e-12
    drift = gt.copy()
    drift[:, 0] += 0.01 * np.arange(200)                 # 1 cm per step along x
    r = E.relative_pose_error(drift, gt)
    assert abs(r["trans_rmse"] - 0.01) < 1e-9 and r["rot_rmse"] < 1e-12
    assert E.absolute_trajectory_error(drift, gt)["rmse"] > 0.1


def test_compare_gt_reports_the_reference_quantities():
    E = _load()
    rd = lambda x, y: {"x": x, "y": y, "theta": 0.0, "range": []}
    out = E.compareGT(rd(1.0, 0.0), rd(0.5, 0.0), rd(2.0, 1.0), rd(1.2, 1.0), rd(10.0, 0.0), rd(9.4, 0.0))
    assert np.allclose(out["trueMove"], (0.6, 0.0, 0.6)) and np.allclose(out["rawMove"], (0.5, 0.0, 0.5))
    assert np.allclose(out["compensateMove"], (0.3, 0.0, 0.3))


def test_reference_trajectory_beats_raw_odometry_on_the_intel_log():
    """The deterministic driver's golden poses (reference output) vs the corrected log: matching must reduce the drift
    of the raw odometry (this pins the metric on real data; the GPU path reproduces these poses bit for bit)."""
    E = _load()
    full, det = load_golden("intel_full.npz"), load_golden("det_intel_full.npz")
    est = det["ref02_poses"].reshape(-1, 3)              # the reference's default geometry (unit 0.02 m), 910 frames
    T = len(est)
    truth, raw = full["truth"][:T], full["poses"][:T]
    a, b = E.evaluate_trajectory(est, truth), E.evaluate_trajectory(raw, truth)
    assert a["frames"] == T == 910
    assert abs(a["ate"]["rmse"] - 0.2518) < 1e-3 and abs(b["ate"]["rmse"] - 24.018) < 1e-2      # known answers
    assert a["ate"]["rmse"] < 0.02 * b["ate"]["rmse"] and a["rpe"]["rot_rmse"] < b["rpe"]["rot_rmse"]

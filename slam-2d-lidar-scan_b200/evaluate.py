"""Accuracy evaluation against ground truth.

The reference only has a debugging printout, ``compareGT`` (Utils/ScanMatcher_OGBased.py:270-289), that compares one
matched step with the corrected log; it never scores a run.  This module keeps that function (same quantities, returned
as a dict, printed on request) and adds the two standard trajectory scores so that runs with different particle counts
can be compared on ``DataSet/PreprocessedData/intel_corrected_log``:

* ATE  -- absolute trajectory error: RMSE / max of the position residuals after the least-squares rigid (SE(2))
          alignment of the estimated onto the true trajectory,
* RPE  -- relative pose error over ``delta`` steps: RMSE of the translational / rotational difference between the
          estimated and the true relative motions (drift per step, independent of the global frame).

numpy only (host-side bookkeeping on [T][3] pose arrays; nothing here is on the hot path).
"""
import math

import numpy as np


def compareGT(currentRawReading, prevRawReading, matchedReading, prevMatchedReading, gtReading, prevGtReading,
              verbose=False):
    """The quantities ScanMatcher_OGBased.py:270-289 prints: true move, raw-odometry move and the matcher's
    compensation on top of the raw move (x, y, r each)."""
    gtMoveX, gtMoveY = gtReading['x'] - prevGtReading['x'], gtReading['y'] - prevGtReading['y']
    rawX, rawY = currentRawReading['x'] - prevRawReading['x'], currentRawReading['y'] - prevRawReading['y']
    compX = matchedReading['x'] - prevMatchedReading['x'] - rawX
    compY = matchedReading['y'] - prevMatchedReading['y'] - rawY
    out = dict(trueMove=(gtMoveX, gtMoveY, math.sqrt(gtMoveX ** 2 + gtMoveY ** 2)),
               rawMove=(rawX, rawY, math.sqrt(rawX ** 2 + rawY ** 2)),
               compensateMove=(compX, compY, math.sqrt(compX ** 2 + compY ** 2)))
    if verbose:
        print("true last pos x: " + str(prevGtReading['x']) + ", y: " + str(prevGtReading['y']))
        print("true curr pos x: " + str(gtReading['x']) + ", y: " + str(gtReading['y']))
        print("true move x: %s, y: %s, r: %s" % out["trueMove"])
        print("Estd last pos x: " + str(prevMatchedReading['x']) + ", y: " + str(prevMatchedReading['y']))
        print("Estd curr pos x: " + str(matchedReading['x']) + ", y: " + str(matchedReading['y']))
        print("raw move x: %s, y: %s, r: %s" % out["rawMove"])
        print("compensate move x: %s, y: %s, r: %s" % out["compensateMove"])
    return out


def align_rigid(est_xy, gt_xy):
    """Least-squares rotation + translation (no scale) taking est_xy [T][2] onto gt_xy -> (R [2][2], t [2])."""
    est_xy, gt_xy = np.asarray(est_xy, dtype=np.float64), np.asarray(gt_xy, dtype=np.float64)
    me, mg = est_xy.mean(0), gt_xy.mean(0)
    H = (est_xy - me).T @ (gt_xy - mg)
    U, _, Vt = np.linalg.svd(H)
    d = np.sign(np.linalg.det(Vt.T @ U.T))
    R = Vt.T @ np.diag([1.0, d]) @ U.T
    return R, mg - R @ me


def absolute_trajectory_error(est, gt):
    """-> dict(rmse, mean, max) of the position residuals [m] after rigid alignment."""
    est, gt = np.asarray(est, dtype=np.float64), np.asarray(gt, dtype=np.float64)
    R, t = align_rigid(est[:, :2], gt[:, :2])
    res = np.linalg.norm((est[:, :2] @ R.T + t) - gt[:, :2], axis=1)
    return dict(rmse=float(np.sqrt(np.mean(res ** 2))), mean=float(res.mean()), max=float(res.max()))


def _wrap(a):
    return (a + np.pi) % (2 * np.pi) - np.pi


def relative_pose_error(est, gt, delta=1):
    """-> dict(trans_rmse [m], rot_rmse [rad]) of the relative-motion differences over ``delta`` steps."""
    est, gt = np.asarray(est, dtype=np.float64), np.asarray(gt, dtype=np.float64)

    def rel(p):
        d = p[delta:, :2] - p[:-delta, :2]
        c, s = np.cos(p[:-delta, 2]), np.sin(p[:-delta, 2])
        return np.stack([c * d[:, 0] + s * d[:, 1], -s * d[:, 0] + c * d[:, 1]], 1), _wrap(p[delta:, 2] - p[:-delta, 2])
    (te, re), (tg, rg) = rel(est), rel(gt)
    dt = np.linalg.norm(te - tg, axis=1)
    dr = _wrap(re - rg)
    return dict(trans_rmse=float(np.sqrt(np.mean(dt ** 2))), rot_rmse=float(np.sqrt(np.mean(dr ** 2))), delta=int(delta))


def evaluate_trajectory(est, gt, delta=1):
    """ATE + RPE of an estimated [T][3] (x, y, theta) trajectory against the ground truth of the same stamps."""
    return dict(frames=int(len(est)), ate=absolute_trajectory_error(est, gt), rpe=relative_pose_error(est, gt, delta))

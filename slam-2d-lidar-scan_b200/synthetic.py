"""Seeded synthetic lidar scenes for benchmarks and large-size property tests.

Rectangular room 16 m x 12 m with 6 random axis-aligned box obstacles, centred on the map origin; the robot drives
a 3 m-radius circle in 0.25 m steps, heading tangent.  Ranges come from exact float64 ray/segment intersection,
rounded to 0.01 m like the Intel log; 5 % of the beams are replaced by the log's max-range sentinel 81.83.
Odometry = true pose + N(0, 0.02 m / 0.01 rad).  numpy only (host-side input generation, not on the hot path).
"""
import numpy as np

SENTINEL = 81.83


def _segments(rng):
    segs = []

    def box(x0, y0, x1, y1):
        segs.extend([(x0, y0, x1, y0), (x1, y0, x1, y1), (x1, y1, x0, y1), (x0, y1, x0, y0)])
    box(-8.0, -6.0, 8.0, 6.0)
    made = 0
    while made < 6:
        w, h = rng.uniform(0.5, 2.0, 2)
        cx, cy = rng.uniform(-7.0, 7.0), rng.uniform(-5.0, 5.0)
        # keep the 3 m driving circle (+ margin) free
        corners = np.array([[cx - w / 2, cy - h / 2], [cx + w / 2, cy - h / 2], [cx - w / 2, cy + h / 2], [cx + w / 2, cy + h / 2]])
        d = np.hypot(corners[:, 0], corners[:, 1])
        if d.min() < 3.6 and np.hypot(cx, cy) + max(w, h) > 2.4:
            continue
        box(cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2)
        made += 1
    return np.array(segs)


def _cast(segs, x, y, angles):
    """Distance along each ray to the nearest segment (inf if none)."""
    dx, dy = np.cos(angles)[:, None], np.sin(angles)[:, None]
    x1, y1, x2, y2 = segs[:, 0][None], segs[:, 1][None], segs[:, 2][None], segs[:, 3][None]
    ex, ey = x2 - x1, y2 - y1
    den = dx * ey - dy * ex
    with np.errstate(divide='ignore', invalid='ignore'):
        t = ((x1 - x) * ey - (y1 - y) * ex) / den
        s = ((x1 - x) * dy - (y1 - y) * dx) / den
    ok = (np.abs(den) > 1e-12) & (t > 1e-9) & (s >= 0.0) & (s <= 1.0)
    t = np.where(ok, t, np.inf)
    return t.min(axis=1)


def make_scene(seed=0, steps=64, K=180, fov=np.pi, unit=0.05, origin=(0.0, 0.0), warm=8, stride=0.25):
    """Returns dict(frames=[reading...], warm=[reading...], truth=[(x,y,theta)...]).

    ``warm`` readings carry TRUE poses snapped to the map lattice (origin + k*unit) and are meant for pre-warming
    the maps with updateOccupancyGrid; ``frames`` carry noisy odometry poses and are fed to the filter.
    ``stride`` is the arc length per step: above 0.3 m the reference's heading prior is active (FastSlam.py:88).
    """
    rng = np.random.default_rng(seed)
    segs = _segments(rng)
    total = warm + steps
    dphi = stride / 3.0
    truth, scans = [], []
    for k in range(total):
        phi = dphi * k
        x, y, th = 3.0 * np.cos(phi), 3.0 * np.sin(phi), phi + np.pi / 2
        ang = np.linspace(th - fov / 2, th + fov / 2, K)
        r = np.round(_cast(segs, x, y, ang), 2)
        r = np.where(np.isfinite(r), r, SENTINEL)
        drop = rng.random(K) < 0.05
        r = np.where(drop, SENTINEL, r)
        truth.append((x + origin[0], y + origin[1], th))
        scans.append(r)
    snap = lambda v, o: o + unit * round((v - o) / unit)
    warmFrames = [dict(x=snap(truth[k][0], origin[0]), y=snap(truth[k][1], origin[1]), theta=truth[k][2],
                       range=scans[k].tolist()) for k in range(warm)]
    frames = []
    odo = np.array(truth[warm])
    odo[0], odo[1] = snap(odo[0], origin[0]), snap(odo[1], origin[1])
    for k in range(warm, total):
        if k > warm:
            step = np.array(truth[k]) - np.array(truth[k - 1])
            odo = odo + step + np.array([rng.normal(0, 0.02), rng.normal(0, 0.02), rng.normal(0, 0.01)])
        frames.append(dict(x=float(odo[0]), y=float(odo[1]), theta=float(odo[2]), range=scans[k].tolist()))
    return dict(frames=frames, warm=warmFrames, truth=truth[warm:], segments=segs)


# BASELINE.json configurations expressed in the reference's constructor arguments
def config(name, origin=(0.0, 0.0)):
    init = {"x": origin[0], "y": origin[1]}
    if name == "c2":      # 128 particles, 500x500 @0.1 m, coarse 11x11x36 / fine 5x5x36
        return dict(N=128, K=180, og=[50, 50, init, 0.1, np.pi, 10, 180, 0.5], sm=[1.1, 0.3, 2, 0.1, 0.25, 0.3, 0.15, 2])
    if name == "c3":      # 1024 particles, 1000x1000 @0.05 m, 10 440 poses
        return dict(N=1024, K=180, og=[50, 50, init, 0.05, np.pi, 10, 180, 0.25], sm=[1.5, 0.3, 2, 0.1, 0.25, 0.3, 0.15, 5])
    if name == "c5":      # stress: 360 beams over 2*pi, 2000x2000 grid
        return dict(N=2048, K=360, og=[100, 100, init, 0.05, 2 * np.pi, 10, 360, 0.25], sm=[1.5, 0.3, 2, 0.1, 0.25, 0.3, 0.15, 5])
    raise KeyError(name)

"""Waypoint generators and the 7-column CSV wire format of the reference's controller/build_traj.py, batched.

Every generator takes `start` (and destinations) either as one 7-vector, like the reference, or as an [N, 7] batch of
per-environment vectors, and returns the trajectory as a torch tensor [T, 7] / [T, N, 7] on `device`, ready to be fed
row by row to `controller.loops.run_trajectory` (the batched `move_j` / `move_l` / `move_l_mug` loops) without a host
round trip.  Same names and argument meaning as the reference (controller/build_traj.py:16-470); differences:

* nothing is written unless `trajectory_fpath` is given (the reference returns the array in that case and writes otherwise:
  here the trajectory is always returned, and also written when a path is given);
* the reference's `build_traj_l_pick_place` adds 0.025 to the caller's `place[2]` in place (`:42`); here the inputs are not
  modified;
* the `*_augmented` noise comes from a torch generator (the reference draws from numpy's global RNG, `:79-111`).

`build_traj_j` / `build_traj_l` reproduce the reference's seeded control points (numpy legacy stream, seeds 42 / 49) and
scipy's `interp1d(kind="cubic")` (a not-a-knot cubic spline) to ~1e-15; tests/golden/build_traj.npz holds the outputs of
the reference's own functions (tools/make_golden_traj.py).
"""
import os

import numpy as np
import torch

COLUMNS_J = ["j1", "j2", "j3", "j4", "j5", "j6", "g"]     # build_traj.py:515
COLUMNS_L = ["x", "y", "z", "rx", "ry", "rz", "g"]


def _t(x, device=None, dtype=torch.float64):
    if isinstance(x, torch.Tensor):
        return x.to(device=device if device is not None else x.device, dtype=dtype)
    return torch.as_tensor(np.asarray(x, dtype=np.float64), dtype=dtype, device=device)


def _fractions(num_points, device, dtype):
    # np.linspace(0, 1, num_points + 1)[1:]: the start point itself is never a waypoint (build_traj.py:215-216)
    return torch.as_tensor(np.linspace(0.0, 1.0, num_points + 1)[1:], dtype=dtype, device=device)


def _finish(traj, trajectory_fpath, ctrl_mode):
    if trajectory_fpath:
        save_traj(traj, trajectory_fpath, ctrl_mode)
    return traj


def build_traj_l_point_custom(start, stop, hold, num_points=15, device=None, dtype=torch.float64):
    """Straight segment start -> stop in `num_points` equal steps, every waypoint held `hold` rows (build_traj.py:187-229)."""
    a, b = _t(start, device, dtype), _t(stop, device, dtype)
    a, b = torch.broadcast_tensors(a, b)
    t = _fractions(num_points, a.device, dtype).reshape((-1,) + (1,) * a.dim())
    pts = (b - a) * t + a                                   # scipy's linear interp1d on the knots (0, 1): slope * x + y0
    return pts.repeat_interleave(int(hold), dim=0)


def build_gripless_traj_gym(start, stop, hold, device=None, dtype=torch.float64):
    """Deprecated 6-column variant with 100 steps (build_traj.py:467-505)."""
    return build_traj_l_point_custom(_t(start, device, dtype)[..., :6], _t(stop, device, dtype)[..., :6], hold, 100, device, dtype)


def build_traj_l_hold(start, hold, trajectory_fpath=None, device=None, dtype=torch.float64):
    """`1000 * hold` copies of the start pose (build_traj.py:16-25)."""
    a = _t(start, device, dtype)
    traj = a.unsqueeze(0).expand((1000 * int(hold),) + tuple(a.shape)).contiguous()
    return _finish(traj, trajectory_fpath, "l")


def _with(v, idx, val):
    out = v.clone()
    out[..., idx] = val
    return out


def build_traj_l_pick_place(start, destinations, hold, trajectory_fpath=None, device=None, dtype=torch.float64):
    """pick -> 0.15 m up with the gripper toggled -> 0.025 m above place -> release (build_traj.py:28-59; `hold` is ignored
    by the reference too: every segment uses 120)."""
    start = _t(start, device, dtype)
    pick, place = (_t(d, start.device, dtype) for d in destinations)
    seg = lambda a, b: build_traj_l_point_custom(a, b, 120, 15, start.device, dtype)
    t_pick = seg(start, pick)
    last = t_pick[-1]
    up = last.clone(); up[..., 2] += 0.15; up[..., 6] += 1 - last[..., 6]
    t_up = seg(last, up)
    above = place.clone(); above[..., 2] += 0.025
    t_place = seg(t_up[-1], above)
    t_drop = seg(t_place[-1], _with(t_place[-1], 6, 0.0))
    return _finish(torch.cat([t_pick, t_up, t_place, t_drop]), trajectory_fpath, "l")


def _imitation_segments(start, block, target):
    """(from, to) pairs of build_traj_l_pick_place_imitation (build_traj.py:126-164), each built from the previous end point."""
    def gen(prev):
        pick = block.clone(); pick[..., 2] = start[..., 2]
        prev = yield prev, pick
        prev = yield prev, _with(prev, 2, block[..., 2])
        prev = yield prev, _with(prev, 6, 1.0)
        up = prev.clone(); up[..., 2] += 0.15
        prev = yield prev, up
        place = target.clone(); place[..., 2] = prev[..., 2]
        prev = yield prev, place
        desc = target.clone(); desc[..., 2] += 0.025
        prev = yield prev, desc
        yield prev, _with(prev, 6, 0.0)
    return gen


def _run_imitation(start, destinations, device, dtype, noise):
    start = _t(start, device, dtype)
    block, target = (_t(d, start.device, dtype) for d in destinations)
    g = _imitation_segments(start, block, target)(start)
    segs, k = [], 0
    frm, to = next(g)
    while True:
        s = build_traj_l_point_custom(frm, to, 100, 15, start.device, dtype)
        if noise is not None:
            s = noise(s, k)
        segs.append(s); k += 1
        try:
            frm, to = g.send(s[-1])
        except StopIteration:
            break
    return torch.cat(segs)


def build_traj_l_pick_place_imitation(start, destinations, hold, trajectory_fpath=None, device=None, dtype=torch.float64):
    """Seven-segment expert demonstration: over the block, down, grab, up, over the target, descend, release
    (build_traj.py:126-164; every segment uses hold = 100, `hold` is ignored by the reference too)."""
    return _finish(_run_imitation(start, destinations, device, dtype, None), trajectory_fpath, "l")


def build_traj_l_pick_place_imitation_augmented(start, destinations, hold, trajectory_fpath=None, device=None, dtype=torch.float64, generator=None):
    """The same demonstration with uniform noise on the waypoints (build_traj.py:60-123): U(0,1)/40 on the first two
    segments, /20 on the others, the last 500 rows of a segment left clean (except the `up` segment, which is noisy throughout)."""
    norms = [40, 40, 20, 20, 20, 20, 20]

    def noise(seg, k):
        n = torch.rand(seg.shape, dtype=dtype, device=seg.device, generator=generator) / norms[k]
        if k != 3:
            n[seg.shape[0] - 500:] = 0
        return seg + n

    return _finish(_run_imitation(start, destinations, device, dtype, noise), trajectory_fpath, "l")


# ---------------------------------------------------------------- seeded cubic trajectories
def _notaknot_cubic(xk, yk, x):
    """Values at x of the not-a-knot cubic spline through (xk, yk) -- what scipy's interp1d(kind='cubic') evaluates.
    xk [K] increasing, yk [..., K], x [P] inside [xk[0], xk[-1]]; float64 numpy (11 knots: solved once on the host)."""
    K = len(xk); h = np.diff(xk)
    A = np.zeros((K, K)); rhs = np.zeros(yk.shape)
    for i in range(1, K - 1):
        A[i, i - 1] = h[i - 1]; A[i, i] = 2 * (h[i - 1] + h[i]); A[i, i + 1] = h[i]
        rhs[..., i] = 6 * ((yk[..., i + 1] - yk[..., i]) / h[i] - (yk[..., i] - yk[..., i - 1]) / h[i - 1])
    # third derivative continuous across the second and the second-to-last knot
    A[0, 0] = h[1]; A[0, 1] = -(h[0] + h[1]); A[0, 2] = h[0]
    A[K - 1, K - 3] = h[K - 2]; A[K - 1, K - 2] = -(h[K - 3] + h[K - 2]); A[K - 1, K - 1] = h[K - 3]
    m = np.linalg.solve(A, rhs[..., None])[..., 0]                    # second derivatives at the knots
    i = np.clip(np.searchsorted(xk, x, side="right") - 1, 0, K - 2)
    hi = h[i]; a = (xk[i + 1] - x) / hi; b = (x - xk[i]) / hi
    return a * yk[..., i] + b * yk[..., i + 1] + ((a ** 3 - a) * m[..., i] + (b ** 3 - b) * m[..., i + 1]) * hi * hi / 6.0


def _seeded_spline(start, bounds, seed, hold, device, dtype):
    start = _t(start, device, dtype)
    s = start.detach().cpu().numpy().astype(np.float64)
    rs = np.random.RandomState(seed)                                   # the stream np.random.seed(seed) starts
    ctrl = np.stack([rs.uniform(lo, hi, 10) for lo, hi in bounds])     # [6, 10], drawn column after column like the reference
    yk = np.concatenate([np.broadcast_to(s[..., :6, None], s.shape[:-1] + (6, 1)), np.broadcast_to(ctrl, s.shape[:-1] + (6, 10))], axis=-1)
    vals = _notaknot_cubic(np.linspace(0, 1, 11), yk, np.linspace(0, 1, 500))        # [..., 6, 500]
    return torch.as_tensor(np.moveaxis(vals, -1, 0), dtype=dtype, device=start.device)   # [500, ..., 6]


def build_traj_j(start, hold, trajectory_fpath=None, device=None, dtype=torch.float64):
    """500 joint-space waypoints on a cubic spline through 10 seeded random control points, gripper open (build_traj.py:383-462)."""
    bounds = [(0.2, 0.5), (-0.3, 0.3), (0.4, 0.8), (0.4, 0.8), (0.4, 0.8), (0.4, 0.8)]
    q = _seeded_spline(start, bounds, 42, hold, device, dtype)
    traj = torch.cat([q, torch.zeros_like(q[..., :1])], dim=-1).repeat_interleave(int(hold), dim=0)
    return _finish(traj, trajectory_fpath, "j")


def build_traj_l(start, hold, trajectory_fpath=None, device=None, dtype=torch.float64):
    """500 Cartesian waypoints on a seeded cubic spline, fixed tool orientation (-1.209, -1.209, 1.209), gripper closed
    (build_traj.py:302-378)."""
    bounds = [(0.2, 0.5), (-0.3, 0.3), (0.4, 0.8), (-0.1, 0.1), (-0.1, 0.1), (-np.pi / 4, np.pi / 4)]
    p = _seeded_spline(start, bounds, 49, hold, device, dtype)
    rot = torch.tensor([-1.209, -1.209, 1.209], dtype=dtype, device=p.device).expand(p.shape[:-1] + (3,))
    traj = torch.cat([p[..., :3], rot, torch.ones_like(p[..., :1])], dim=-1).repeat_interleave(int(hold), dim=0)
    return _finish(traj, trajectory_fpath, "l")


def build_traj_l_pick_move_place(start, destinations, hold, trajectory_fpath=None, device=None, dtype=torch.float64):
    """pick (hold) -> seeded spline excursion (120) -> place (50) -> release (5) (build_traj.py:167-199)."""
    start = _t(start, device, dtype)
    pick, place = (_t(d, start.device, dtype) for d in destinations)
    t_pick = build_traj_l_point_custom(start, pick, hold, 15, start.device, dtype)
    t_move = build_traj_l(t_pick[-1], 120, None, start.device, dtype)
    t_place = build_traj_l_point_custom(t_move[-1], place, 50, 15, start.device, dtype)
    t_drop = build_traj_l_point_custom(t_place[-1], _with(t_place[-1], 6, 0.0), 5, 15, start.device, dtype)
    return _finish(torch.cat([t_pick, t_move, t_place, t_drop]), trajectory_fpath, "l")


# ---------------------------------------------------------------- CSV wire format
def save_traj(traj, trajectory_fpath, ctrl_mode):
    """Header `j1..j6,g` / `x,y,z,rx,ry,rz,g`, one row per step, shortest round-trip float text like pandas' to_csv
    (build_traj.py:509-525).  Only single-environment trajectories [T, 7] have a file form."""
    a = traj.detach().cpu().numpy() if isinstance(traj, torch.Tensor) else np.asarray(traj)
    if a.ndim != 2 or a.shape[1] != 7:
        raise ValueError("save_traj expects a [T, 7] trajectory, got %s" % (a.shape,))
    cols = COLUMNS_J if ctrl_mode == "j" else COLUMNS_L
    d = os.path.dirname(trajectory_fpath)
    if d:
        os.makedirs(d, exist_ok=True)
    with open(trajectory_fpath, "w") as f:
        f.write(",".join(cols) + "\n")
        for row in a.astype(np.float64):
            f.write(",".join(repr(float(v)) for v in row) + "\n")


def load_trajectory(trajectory_fpath, device=None, dtype=torch.float64):
    """The reference's loader (controller/aux.py:107-111: genfromtxt, one header row, reshape(-1, 7)) as a torch tensor."""
    a = np.genfromtxt(trajectory_fpath, delimiter=",", skip_header=1).reshape(-1, 7)
    return torch.as_tensor(a, dtype=dtype, device=device)


load_trajectory_file = load_trajectory   # name used by the reference's archived scripts

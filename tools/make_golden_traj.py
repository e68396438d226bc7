#!/usr/bin/env python3
"""Generate tests/golden/build_traj.npz + build_traj_l.csv by running the REFERENCE's own controller/build_traj.py
(from /root/reference, on top of oracle/refshim.py for its `mujoco` import).  Build container only:

    python tools/make_golden_traj.py

Pins ur3e_b200/controller/build_traj.py: the waypoint generators (linear point-to-point segments with `hold`
repeats, the scripted pick-and-place sequences, the seeded cubic-spline joint / Cartesian trajectories incl.
numpy's legacy RNG stream and scipy's not-a-knot cubic interp1d) and the 7-column CSV wire format.
"""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")

from oracle import refshim  # noqa: E402

refshim.import_reference()
from controller import build_traj as BT  # noqa: E402  (the reference module)

rng = np.random.default_rng(7)
start = np.array([0.3, 0.13, 0.17, -1.209, -1.209, 1.209, 0.0])
pick = np.array([0.32, 0.05, 0.06, -1.209, -1.209, 1.209, 1.0])
place = np.array([0.30, 0.25, 0.08, -1.209, -1.209, 1.209, 1.0])
rec = dict(start=start, pick=pick, place=place)
rec["point_custom"] = BT.build_traj_l_point_custom(start, pick, hold=3)
rec["point_custom_30"] = BT.build_traj_l_point_custom(start, place, hold=2, num_points=30)
rec["hold"] = BT.build_traj_l_hold(start, hold=1)[::250]          # 1000 identical rows; keep every 250th
rec["pick_place"] = BT.build_traj_l_pick_place(start, [pick.copy(), place.copy()], hold=120)[::60]
rec["pick_place_len"] = np.int64(BT.build_traj_l_pick_place(start, [pick.copy(), place.copy()], hold=120).shape[0])
rec["imitation"] = BT.build_traj_l_pick_place_imitation(start, [pick.copy(), place.copy()], hold=100)[::50]
rec["traj_j"] = BT.build_traj_j(np.zeros(7), hold=2)
rec["traj_l"] = BT.build_traj_l(start, hold=1)
rec["gripless"] = BT.build_gripless_traj_gym(start[:6], pick[:6], hold=2)
# batch of random segments for the vectorised path
S = rng.uniform(-0.5, 0.5, (5, 7)); E = rng.uniform(-0.5, 0.5, (5, 7))
rec["batch_start"] = S; rec["batch_stop"] = E
rec["batch_segments"] = np.stack([BT.build_traj_l_point_custom(S[i], E[i], hold=4, num_points=10) for i in range(5)], axis=1)
np.savez_compressed(os.path.join(OUT, "build_traj.npz"), **rec)
# CSV wire format (save_traj writes relative to the cwd and prints)
cwd = os.getcwd(); os.chdir("/tmp")
with contextlib.redirect_stdout(io.StringIO()):
    BT.save_traj(rec["point_custom"][::3], "/tmp/golden_traj_l.csv", ctrl_mode="l")
    BT.save_traj(rec["traj_j"][:6], "/tmp/golden_traj_j.csv", ctrl_mode="j")
os.chdir(cwd)
for n in ("l", "j"):
    with open("/tmp/golden_traj_%s.csv" % n) as f, open(os.path.join(OUT, "build_traj_%s.csv" % n), "w") as g:
        g.write(f.read())
print({k: getattr(v, "shape", v) for k, v in rec.items()})

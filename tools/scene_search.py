"""Scene-parameter search (CPU, this container): run the REFERENCE'S OWN tracking loop (oracle/_ref) over a
rendered synthetic sequence and report how far the estimated trajectory is from ground truth across keyframe
switches.  Test infrastructure; not part of the product path.

    python tools/scene_search.py --frames 100 --seq 3 --bands street      (or: isotropic, or a JSON band list)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(args):
    import torch
    torch.set_num_threads(args.threads)
    from odometry_b200 import synth
    from oracle import ref

    kw = {}
    for k in ("wall_x", "ground_y"):
        v = getattr(args, k)
        if v is not None:
            kw[k] = v
    if args.bands == "isotropic":
        kw["bands"] = synth.ISOTROPIC_BANDS
    elif args.bands != "street":
        kw["bands"] = tuple(tuple(b) for b in json.loads(args.bands))
    sc = synth.Scene(seed=synth.BASE_SEED + args.seq, end_wall_z=args.end_wall + args.step * args.frames, **kw)
    poses = synth.trajectory(args.frames, step=args.step, yaw_amp_deg=args.yaw)
    trk = ref.Tracker(1241, 376, 4)
    out = []
    t0 = time.time()
    n_kf = 0
    for f, T in enumerate(poses):
        L, R, _ = synth.render_stereo(sc, T, 1241, 376, "cpu")
        rc, pa, pk, kf, mag = trk.step(L.numpy().astype(np.float32), R.numpy().astype(np.float32))
        n_kf += int(kf)
        gt = np.linalg.inv(poses[0]) @ T
        err_t = float(np.linalg.norm(pa[:3, 3] - gt[:3, 3]))
        dist = float(np.linalg.norm(gt[:3, 3]))
        out.append(dict(f=f, rc=rc, kf=bool(kf), mag=round(float(mag), 4), err_t=round(err_t, 4),
                        rel=round(err_t / max(dist, 1e-9), 5), z=round(float(pa[2, 3]), 3), zgt=round(float(gt[2, 3]), 3)))
        if args.verbose:
            print(out[-1], flush=True)
        if rc != 0:
            break
    rels = [o["rel"] for o in out[1:]]
    post = [out[i]["err_t"] for i in range(2, len(out)) if out[i - 1]["kf"]]
    rest = [out[i]["err_t"] for i in range(1, len(out)) if not out[i - 1]["kf"]]
    res = dict(args={k: getattr(args, k) for k in ("seq", "frames", "step", "yaw", "bands", "wall_x", "ground_y")},
               n_kf=n_kf, max_rel=max(rels) if rels else None, final_rel=rels[-1] if rels else None,
               final_err=out[-1]["err_t"], max_err=max(o["err_t"] for o in out),
               post_switch_mean=round(float(np.mean(post)), 4) if post else None, post_switch_max=max(post) if post else None,
               other_max=max(rest) if rest else None, secs=round(time.time() - t0, 1))
    print(json.dumps(res), flush=True)
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seq", type=int, default=0)
    ap.add_argument("--frames", type=int, default=30)
    ap.add_argument("--step", type=float, default=0.8)
    ap.add_argument("--yaw", type=float, default=0.4)
    ap.add_argument("--bands", default="street", help="street | isotropic | JSON list of synth band tuples")
    ap.add_argument("--wall-x", dest="wall_x", type=float, default=None)
    ap.add_argument("--ground-y", dest="ground_y", type=float, default=None)
    ap.add_argument("--end-wall", dest="end_wall", type=float, default=60.0)
    ap.add_argument("--threads", type=int, default=2)
    ap.add_argument("--verbose", action="store_true")
    run(ap.parse_args())

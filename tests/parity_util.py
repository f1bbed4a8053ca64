"""Helpers shared by the parity tests: drive the reference oracle (oracle/refapi.py, TEST ONLY) and the CUDA
library through the same calls and compare."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

EUROC_CFG = dict(name="euroc", cam=dict(w=752, h=480, zfx=458.654, zfy=457.296, ppx=367.215, ppy=248.375),
                 sigma0=3.56359, ksigma=1.2599, thresh=0.01, gain=5e-7, tmax=0.5, tmin=0.005, kl_max=40000,
                 kl_ref=15000, track_points=12000, radius=40, match_thresh=0.5, iter_max=5, init_type=2,
                 init_iter=2, reweight=2.0, match_num_thresh=0, thr_mod=1.0, thr_ang=45.0, loc_unc_match=2.0,
                 reg_thresh=0.5, q_abs=1e-4, loc_unc=1.0)
TUM_CFG = dict(name="tum", cam=dict(w=640, h=480, zfx=525.0, zfy=525.0, ppx=320.0, ppy=240.0),
               sigma0=1.7818, ksigma=1.2599, thresh=0.01, gain=0.0, tmax=0.05, tmin=0.03, kl_max=25000,
               kl_ref=15000, track_points=12000, radius=20, match_thresh=1.0, iter_max=10, init_type=2,
               init_iter=2, reweight=2.0, match_num_thresh=4, thr_mod=1.0, thr_ang=45.0, loc_unc_match=2.0,
               reg_thresh=0.5, q_abs=1e-4, loc_unc=1.0)
# BASELINE.json configs[3]: synthetic 1280x960, ~30 k keylines
BIG_CFG = dict(name="big", cam=dict(w=1280, h=960, zfx=780.0, zfy=778.0, ppx=640.5, ppy=479.25),
               sigma0=3.56359, ksigma=1.2599, thresh=0.01, gain=5e-7, tmax=0.5, tmin=0.005, kl_max=40000,
               kl_ref=30000, track_points=24000, radius=40, match_thresh=0.5, iter_max=5, init_type=2,
               init_iter=2, reweight=2.0, match_num_thresh=0, thr_mod=1.0, thr_ang=45.0, loc_unc_match=2.0,
               reg_thresh=0.5, q_abs=1e-4, loc_unc=1.0)
POS_NEG, DOG_THRESH, PLANE_FIT = 0.4, 0.095259868922420, 2

# KeyLine fields the reference defines at every stage (m_m0 / n_m0 / score are uninitialised until a match)
KL_EXACT_DETECT = ["p_inx", "m_m", "u_m", "n_m", "c_p", "rho", "s_rho", "rho0", "s_rho0", "p_m", "p_m_0",
                   "m_id", "m_id_f", "m_num", "p_id", "n_id"]


class Report:
    def __init__(self, name):
        self.name = name
        self.items = []

    def add(self, what, ok, info=""):
        self.items.append(dict(what=what, ok=bool(ok), info=str(info)))
        print("[%s] %-46s %s %s" % (self.name, what, "PASS" if ok else "FAIL", info))

    def exact(self, what, a, b):
        a, b = np.asarray(a), np.asarray(b)
        if a.shape != b.shape:
            self.add(what, False, "shape %s vs %s" % (a.shape, b.shape))
            return False
        if a.dtype.kind == "f":
            same = (a.view(np.uint32 if a.dtype.itemsize == 4 else np.uint64)
                    == b.view(np.uint32 if b.dtype.itemsize == 4 else np.uint64)) | (np.isnan(a) & np.isnan(b))
        else:
            same = a == b
        nbad = int(same.size - same.sum())
        info = "bitwise n=%d" % same.size
        if nbad:
            idx = np.argwhere(~same)[:3]
            info = "mismatch %d/%d first=%s ref=%s got=%s" % (nbad, same.size, idx.tolist(),
                                                               [a[tuple(i)] for i in idx], [b[tuple(i)] for i in idx])
        self.add(what, nbad == 0, info)
        return nbad == 0

    def close(self, what, ref, got, rtol, atol=0.0):
        ref, got = np.asarray(ref, np.float64), np.asarray(got, np.float64)
        if ref.shape != got.shape:
            self.add(what, False, "shape %s vs %s" % (ref.shape, got.shape))
            return False
        both_nan = np.isnan(ref) & np.isnan(got)
        err = np.abs(ref - got)
        tol = atol + rtol * np.abs(ref)
        bad = ~((err <= tol) | both_nan)
        nbad = int(bad.sum())
        mx = float(np.nanmax(err / (np.abs(ref) + atol + 1e-300))) if err.size else 0.0
        self.add(what, nbad == 0, "max rel err %.3e (rtol %.1e atol %.1e) bad=%d/%d" % (mx, rtol, atol, nbad, err.size))
        return nbad == 0

    def ok(self):
        return all(i["ok"] for i in self.items)

    def dump(self):
        d = os.path.join(ROOT, "gpurun_out")
        try:
            os.makedirs(d, exist_ok=True)
            with open(os.path.join(d, "parity_%s.json" % self.name), "w") as f:
                json.dump(self.items, f, indent=1)
        except OSError:
            pass

    def failures(self):
        return [i for i in self.items if not i["ok"]]


def compare_keylines(rep, tag, ref, got, exact_fields, close_fields=(), rtol=1e-12):
    if len(ref) != len(got):
        rep.add(tag + " kn", False, "%d vs %d" % (len(ref), len(got)))
        return
    rep.add(tag + " kn", True, "kn=%d" % len(ref))
    for f in exact_fields:
        rep.exact("%s %s" % (tag, f), ref[f], got[f])
    for f in close_fields:
        rep.close("%s %s" % (tag, f), ref[f], got[f], rtol)

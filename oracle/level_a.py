"""Level-A CPU baseline: the reference's own functions, compiled unmodified into
oracle/_ref/libref_mtrack.so, timed ONE THREAD, stage by stage, on one frame pair.  TEST / MEASUREMENT INFRASTRUCTURE ONLY
(bench.py's cpu_baseline leg and tools/); the product never imports this.

Each repetition runs the whole per-frame chain of the reference's first and second thread on a fresh pair of edge maps
(rebvo_first_t.cpp:213-272, rebvo_second_t.cpp:167-585), so every stage sees the state the previous one left; the table is
the median over the repetitions."""
import time

import numpy as np

from . import refapi

POS_NEG, DOG_THRESH, PLANE_FIT = 0.4, 0.095259868922420, 2

EUROC = dict(name="752x480 EuRoC parameters", cam=dict(w=752, h=480, zfx=458.654, zfy=457.296, ppx=367.215, ppy=248.375),
             sigma0=3.56359, ksigma=1.2599, thresh=0.01, gain=5e-7, tmax=0.5, tmin=0.005, kl_max=40000, kl_ref=15000,
             track_points=12000, radius=40, match_thresh=0.5, iter_max=5, init_type=2, init_iter=2, reweight=2.0,
             match_num_thresh=0, thr_mod=1.0, thr_ang=45.0, loc_unc_match=2.0, reg_thresh=0.5, q_abs=1e-4, loc_unc=1.0)
TUM = dict(name="640x480 TUM desk parameters", cam=dict(w=640, h=480, zfx=525.0, zfy=525.0, ppx=320.0, ppy=240.0),
           sigma0=1.7818, ksigma=1.2599, thresh=0.01, gain=0.0, tmax=0.05, tmin=0.03, kl_max=25000, kl_ref=15000,
           track_points=12000, radius=20, match_thresh=1.0, iter_max=10, init_type=2, init_iter=2, reweight=2.0,
           match_num_thresh=4, thr_mod=1.0, thr_ang=45.0, loc_unc_match=2.0, reg_thresh=0.5, q_abs=1e-4, loc_unc=1.0)
BIG = dict(name="1280x960, 30 k keylines", cam=dict(w=1280, h=960, zfx=780.0, zfy=778.0, ppx=640.5, ppy=479.25),
           sigma0=3.56359, ksigma=1.2599, thresh=0.01, gain=5e-7, tmax=0.5, tmin=0.005, kl_max=40000, kl_ref=30000,
           track_points=24000, radius=40, match_thresh=0.5, iter_max=5, init_type=2, init_iter=2, reweight=2.0,
           match_num_thresh=0, thr_mod=1.0, thr_ang=45.0, loc_unc_match=2.0, reg_thresh=0.5, q_abs=1e-4, loc_unc=1.0)

STAGES = ["ConvertRGB2BW", "sspace::build", "edge_finder::detect", "reEstimateThresh + EstimateQuantile",
          "global_tracker::build_field", "Minimizer_RV", "FordwardMatch + rotate_keylines", "directed_matching",
          "Regularize_1_iter", "UpdateInverseDepthKalman", "EstimateReScalingOpt"]


def _seed_depth(kl):
    kl = kl.copy()
    kl["rho"] = 1.0
    kl["s_rho"] = 0.2
    kl["m_num"] = 5
    return kl


def stage_table(cfg, f0, f1, reps=5, thresh=None):
    """{stage: median ms} of the reference's per-frame chain, single thread; plus `kn` (keylines of the new frame) and
    `sum`.  `thresh`: detector threshold to start from (default: run two settling passes of the auto-threshold)."""
    cam = cfg["cam"]
    acc = {s: [] for s in STAGES}
    kn = 0
    t_start = cfg["thresh"] if thresh is None else thresh
    for rep in range(reps + 1):   # (rep 0 = warm-up, also settles the threshold)
        maps = [refapi.RefMap(cam["w"], cam["h"], cam["ppx"], cam["ppy"], cam["zfx"], cam["zfy"], cfg["sigma0"], cfg["ksigma"])
                for _ in range(2)]
        t = {s: 0.0 for s in STAGES}
        th, lk = t_start, 0
        rt = 0.0
        for i, fr in enumerate((f0, f1)):
            m = maps[i]
            c0 = time.perf_counter()
            m.rgb2bw(fr)
            c1 = time.perf_counter()
            m.build()
            c2 = time.perf_counter()
            kn_i, th, lk = m.detect(PLANE_FIT, POS_NEG, DOG_THRESH, cfg["kl_max"], th, lk, cfg["kl_ref"], cfg["gain"],
                                    cfg["tmax"], cfg["tmin"])
            c3 = time.perf_counter()
            _, rt = m.reestimate(cfg["track_points"], 100)
            c4 = time.perf_counter()
            if i == 1:   # the new frame's detector stages
                t["ConvertRGB2BW"] = c1 - c0
                t["sspace::build"] = c2 - c1
                t["edge_finder::detect"] = c3 - c2
                t["reEstimateThresh + EstimateQuantile"] = c4 - c3
                kn = kn_i
        if rep == 0:
            t_start = th
        old, new = maps
        old.set_keylines(_seed_depth(old.keylines()))
        c0 = time.perf_counter()
        q = old.quantile(1e-3, 20.0, 0.9, 100)
        c1 = time.perf_counter()
        new.build_field(cfg["radius"], rt)
        c2 = time.perf_counter()
        mr = new.minimizer_rv(old, np.zeros(3), np.zeros(3), cfg["match_thresh"], cfg["iter_max"], cfg["init_type"],
                              cfg["reweight"], q, cfg["match_num_thresh"], cfg["init_iter"])
        c3 = time.perf_counter()
        old.forward_match(new)
        R0 = refapi.so3_exp(mr["W"])
        old.rotate(R0)
        c4 = time.perf_counter()
        R = R0.T.copy()
        new.directed_matching(old, mr["V"], mr["RVel"], R, cfg["thr_mod"], cfg["thr_ang"], cfg["radius"], cfg["loc_unc_match"])
        c5 = time.perf_counter()
        new.regularize(cfg["reg_thresh"])
        c6 = time.perf_counter()
        new.ekf(mr["V"], mr["RVel"], mr["RW0"], cfg["q_abs"], 1.6968e-4, cfg["loc_unc"])
        c7 = time.perf_counter()
        new.rescale(20.0, 1, False)
        c8 = time.perf_counter()
        t["reEstimateThresh + EstimateQuantile"] += c1 - c0
        t["global_tracker::build_field"] = c2 - c1
        t["Minimizer_RV"] = c3 - c2
        t["FordwardMatch + rotate_keylines"] = c4 - c3
        t["directed_matching"] = c5 - c4
        t["Regularize_1_iter"] = c6 - c5
        t["UpdateInverseDepthKalman"] = c7 - c6
        t["EstimateReScalingOpt"] = c8 - c7
        if rep > 0:
            for s in STAGES:
                acc[s].append(t[s] * 1e3)
        del maps
    out = {s: float(np.median(acc[s])) for s in STAGES}
    out["sum"] = float(sum(out[s] for s in STAGES))
    out["kn"] = int(kn)
    return out

"""TEST INFRASTRUCTURE ONLY: ctypes binding of the CPU restatement (oracle/rebvo_oracle.cpp).  Same method names
as oracle/refapi.RefMap so that tests can run one flow against either."""
import ctypes as C
import os

import numpy as np

from . import build_port
from .refapi import KEYLINE

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_port.build())
        L.orc_map_create.restype = C.c_void_p
        L.orc_map_create.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_double,
                                     C.c_double]
        for f in ("orc_quantile", "orc_minimizer_rv", "orc_try_vel_rot", "orc_rescale", "orc_try_vel",
                  "orc_minimizer_v"):
            getattr(L, f).restype = C.c_double
        L.orc_reestimate.restype = C.c_float
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class PortMap:
    def __init__(self, w, h, ppx, ppy, zfx, zfy, sigma0, ksigma):
        self.w, self.h = w, h
        self.L = lib()
        self.h_ = C.c_void_p(self.L.orc_map_create(w, h, ppx, ppy, zfx, zfy, sigma0, ksigma))

    def __del__(self):
        if getattr(self, "h_", None):
            self.L.orc_map_destroy(self.h_)
            self.h_ = None

    def box_plan(self):
        d, s = np.zeros(6, np.int32), np.zeros(2)
        self.L.orc_box_plan(self.h_, _p(d), _p(s))
        return d.reshape(2, 3), s

    def rgb2bw(self, rgb):
        rgb = np.ascontiguousarray(rgb, np.uint8)
        self.L.orc_rgb2bw(self.h_, _p(rgb))

    def build(self):
        self.L.orc_build(self.h_)

    def plane(self, which):
        idx = {"img0": 0, "img1": 1, "dog": 2, "dx": 3, "dy": 4, "gray": 5}[which]
        out = np.empty((self.h, self.w), np.float32)
        self.L.orc_get_plane(self.h_, idx, _p(out))
        return out

    def detect(self, plane_fit, pos_neg, dog_thresh, kl_max, tresh, l_kl_num, kl_ref, gain, tmax, tmin):
        t, l = C.c_double(tresh), C.c_int(l_kl_num)
        kn = self.L.orc_detect(self.h_, plane_fit, C.c_double(pos_neg), C.c_double(dog_thresh), kl_max, C.byref(t),
                               C.byref(l), kl_ref, C.c_double(gain), C.c_double(tmax), C.c_double(tmin))
        return kn, t.value, l.value

    def reestimate(self, knum, n):
        v = self.L.orc_reestimate(self.h_, knum, n)
        return int(v), v

    def knum(self):
        return self.L.orc_knum(self.h_)

    def keylines(self):
        out = np.zeros(self.knum(), KEYLINE)
        self.L.orc_get_keylines(self.h_, _p(out))
        return out

    def set_keylines(self, kl):
        kl = np.ascontiguousarray(kl, KEYLINE)
        self.L.orc_set_keylines(self.h_, _p(kl), len(kl))

    def mask(self):
        out = np.empty((self.h, self.w), np.int32)
        self.L.orc_get_mask(self.h_, _p(out))
        return out

    def set_mask(self, mask, kn=None):
        mask = np.ascontiguousarray(mask, np.int32)
        self.L.orc_set_mask(self.h_, _p(mask))

    def quantile(self, smin, smax, perc, n):
        return self.L.orc_quantile(self.h_, C.c_double(smin), C.c_double(smax), C.c_double(perc), n)

    def build_field(self, radius, min_mod):
        self.L.orc_build_field(self.h_, radius, C.c_float(min_mod))

    def field(self):
        out = np.empty((self.h, self.w, 2), np.int32)
        self.L.orc_get_field(self.h_, _p(out))
        return out

    def set_frame_count(self, fc):
        self.L.orc_set_frame_count(self.h_, C.c_uint(fc))

    def try_vel_rot(self, old, X, reweight, procjf, match_thresh, s_rho_min, match_num_thresh, k_huber, res_in):
        X = np.array(X, np.float64)
        pnum = (old.knum() + 3) & ~3
        res_in = np.ascontiguousarray(res_in, np.float64)
        assert len(res_in) >= old.knum()
        res_out = np.full(pnum, np.nan)
        JtJ, JtF = np.zeros((6, 6)), np.zeros(6)
        s = self.L.orc_try_vel_rot(self.h_, old.h_, _p(X), int(reweight), int(procjf), C.c_double(match_thresh),
                                   C.c_double(s_rho_min), C.c_uint(match_num_thresh), C.c_double(k_huber), _p(res_in),
                                   _p(res_out), _p(JtJ), _p(JtF))
        return s, JtJ, JtF, res_out

    # ---- IMU-mode rows (same signatures as oracle/refapi.py) ----------------------------------------------
    def try_vel(self, old, V, match_thresh, s_rho_min, match_num_thresh, residuals, rw_dist, min_mod):
        V = np.array(V, np.float64)
        res = np.ascontiguousarray(residuals[:old.knum()], np.float64).copy()
        JtJ, JtF = np.zeros((3, 3)), np.zeros(3)
        s = self.L.orc_try_vel(self.h_, old.h_, _p(V), C.c_double(match_thresh), C.c_double(s_rho_min),
                               C.c_uint(match_num_thresh), _p(res), C.c_double(rw_dist), C.c_float(min_mod), _p(JtJ),
                               _p(JtF))
        return s, JtJ, JtF, res

    def minimizer_v(self, old, V, match_thresh, iter_max, s_rho_min, match_num_thresh, rw_dist, min_mod):
        V = np.array(V, np.float64)
        RV = np.zeros((3, 3))
        F = self.L.orc_minimizer_v(self.h_, old.h_, _p(V), _p(RV), C.c_double(match_thresh), iter_max,
                                   C.c_double(s_rho_min), C.c_uint(match_num_thresh), C.c_double(rw_dist),
                                   C.c_float(min_mod))
        return dict(F=F, V=V, RVel=RV)

    def ext_rot_vel(self, V, loc_unc, hub):
        V = np.array(V, np.float64)
        Wx, Rx, X = np.zeros((6, 6)), np.zeros((6, 6)), np.zeros(6)
        ok = self.L.orc_ext_rot_vel(self.h_, _p(V), _p(Wx), _p(Rx), _p(X), C.c_double(loc_unc), C.c_double(hub))
        return bool(ok), Wx, Rx, X

    def minimizer_rv(self, old, V, W, match_thresh, iter_max, init_type, reweight, max_s_rho, match_num_thresh,
                     init_iter):
        V, W = np.array(V, np.float64), np.array(W, np.float64)
        RV, RW, WX = np.zeros((3, 3)), np.zeros((3, 3)), np.zeros((6, 6))
        e1, e2 = C.c_double(0), C.c_double(0)
        F = self.L.orc_minimizer_rv(self.h_, old.h_, _p(V), _p(W), _p(RV), _p(RW), C.c_double(match_thresh), iter_max,
                                    init_type, C.c_double(reweight), C.byref(e1), C.byref(e2), C.c_double(max_s_rho),
                                    C.c_uint(match_num_thresh), C.c_double(init_iter), _p(WX))
        return dict(F=F, V=V, W=W, RVel=RV, RW0=RW, W_X=WX, rel_err=e1.value, rel_err_score=e2.value)

    def forward_match(self, new):
        return self.L.orc_forward_match(self.h_, new.h_)

    def rotate(self, R):
        R = np.ascontiguousarray(R, np.float64)
        self.L.orc_rotate(self.h_, _p(R))

    def directed_matching(self, old, V, RVel, BackRot, thr_mod, thr_ang, max_radius, loc_unc):
        V = np.ascontiguousarray(V, np.float64)
        RVel = np.ascontiguousarray(RVel, np.float64)
        BackRot = np.ascontiguousarray(BackRot, np.float64)
        return self.L.orc_directed_matching(self.h_, old.h_, _p(V), _p(RVel), _p(BackRot), C.c_double(thr_mod),
                                            C.c_double(thr_ang), C.c_double(max_radius), C.c_double(loc_unc))

    def regularize(self, thresh):
        return self.L.orc_regularize(self.h_, C.c_double(thresh))

    def ekf(self, V, RVel, RW0, qabs, qrel, loc_unc):
        V = np.ascontiguousarray(V, np.float64)
        self.L.orc_ekf(self.h_, _p(V), C.c_double(qabs), C.c_double(loc_unc))

    def rescale(self, s_rho_min, match_num_min, re_escale):
        rkp = C.c_double(0)
        kp = self.L.orc_rescale(self.h_, C.byref(rkp), C.c_double(s_rho_min), C.c_uint(match_num_min), int(re_escale))
        return kp, rkp.value


def so3_exp(w):
    w = np.ascontiguousarray(w, np.float64)
    R = np.zeros((3, 3))
    lib().orc_so3_exp(_p(w), _p(R))
    return R


def bias_correct(X, Wx, Gb, Wb, Rg, Rb):
    """edge_tracker::BiasCorrect restated (in/out arrays like the reference)."""
    a = [np.ascontiguousarray(np.array(v, np.float64)) for v in (X, Wx, Gb, Wb)]
    Rg, Rb = np.ascontiguousarray(Rg, np.float64), np.ascontiguousarray(Rb, np.float64)
    lib().orc_bias_correct(_p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(Rg), _p(Rb))
    return a


def undistort_rgb(cam, kc, rgb):
    """image_undistort(cam).undistort<true>(out, in) restated, one RGB24 frame."""
    rgb = np.ascontiguousarray(rgb, np.uint8)
    out = np.zeros_like(rgb)
    kc = np.ascontiguousarray(kc, np.float64)
    lib().orc_undistort_rgb(cam["w"], cam["h"], C.c_float(cam["ppx"]), C.c_float(cam["ppy"]), C.c_float(cam["zfx"]),
                            C.c_float(cam["zfy"]), _p(kc), _p(rgb), _p(out))
    return out


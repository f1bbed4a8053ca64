"""Host filter chain of the IMU mode (rebvo_b200/csrc/imu_filter.h) against the reference's own ScaleEstimator and
ImuGrabber (unmodified sources in oracle/_ref/libref_mtrack.so): inter-frame gyro integration, EstAcelLsq4, MeanAcel4,
estKaGMEKBias (20 Gauss-Newton steps on the 7-state scale / gravity / bias posterior).  Runs on the CPU."""
import ctypes as C
import os

import numpy as np
import pytest


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def libs(built):
    from oracle import refapi
    from rebvo_b200 import capi
    if not refapi.available():
        pytest.skip("compiled reference not available")
    R = refapi.lib()
    if not hasattr(R, "ref_est_ka_gmek_bias"):
        pytest.skip("reference library without the IMU exports")
    L = C.CDLL(capi.LIB_PATH)
    L.rb_hostmath_imu_hist_new.restype = C.c_void_p
    L.rb_hostmath_est_ka_gmek_bias.restype = C.c_double
    R.ref_est_ka_gmek_bias.restype = C.c_double
    return L, R


def _rot(rng, s=0.02):
    from oracle import refapi
    return refapi.so3_exp(rng.normal(0, s, 3))


def test_imu_integration(libs):
    from rebvo_b200 import synth
    L, R = libs
    seq = synth.Sequence(w=64, h=48, seed=1)
    smp = np.ascontiguousarray(synth.imu_samples(seq, 40, gyro_noise=2e-3))
    ts = np.arange(40) / 20.0
    a, b = np.zeros((40, 20)), np.zeros((40, 20))
    L.rb_hostmath_imu_integrate(_p(smp), len(smp), _p(ts), 40, _p(a))
    R.ref_imu_integrate(_p(smp), len(smp), _p(ts), 40, _p(b))
    assert np.array_equal(a[:, 0], b[:, 0]) and (a[:, 0] > 0).all()          # sample counts per interval
    assert np.allclose(a, b, rtol=1e-13, atol=1e-15)


def test_acceleration_histories(libs):
    L, R = libs
    rng = np.random.default_rng(3)
    h = C.c_void_p(L.rb_hostmath_imu_hist_new())
    acc_a, acc_b = np.zeros(3), np.zeros(3)
    for k in range(12):   # the reference keeps its history in function statics: one sequence per process
        vel, sa, Rm = rng.normal(0, 0.3, 3), rng.normal(0, 5, 3), np.ascontiguousarray(_rot(rng))
        dt = float(rng.uniform(0.03, 0.07))
        L.rb_hostmath_est_acel_lsq4(h, _p(vel), _p(acc_a), _p(Rm), C.c_double(dt))
        R.ref_est_acel_lsq4(_p(vel), _p(acc_b), _p(Rm), C.c_double(dt))
        assert np.allclose(acc_a, acc_b, rtol=1e-9, atol=1e-12), (k, acc_a, acc_b)
        ma, mb = np.zeros(3), np.zeros(3)
        L.rb_hostmath_mean_acel4(h, _p(sa), _p(ma), _p(Rm))
        R.ref_mean_acel4(_p(sa), _p(mb), _p(Rm))
        assert np.allclose(ma, mb, rtol=1e-14, atol=1e-15)
    L.rb_hostmath_imu_hist_free(h)


def test_scale_gravity_bias_filter(libs):
    L, R = libs
    rng = np.random.default_rng(5)
    g = 9.8
    Xa = np.array([np.pi / 4, 0.1, g, -0.2, 0, 0, 0.0])
    Pa = np.diag([1.2e-3 ** 2, 100, 100, 100, 1e-13, 1e-13, 1e-13]).astype(np.float64)
    Xb, Pb = Xa.copy(), Pa.copy()
    Qg, Qbias = np.eye(3) * 2e-3 ** 2, np.eye(3) * 1e-14
    Rs = np.eye(3) * 2e-3 ** 2
    for k in range(8):
        a_v = rng.normal(0, 0.5, 3)
        scale = 1.3
        a_s = scale * a_v - Xa[1:4] + rng.normal(0, 2e-3, 3)
        Rot = np.ascontiguousarray(_rot(rng, 5e-3))
        A = rng.normal(0, 1, (6, 6))
        Wvw = np.ascontiguousarray(A @ A.T + np.eye(6) * 50) * 1e4
        Qrot = np.ascontiguousarray(np.linalg.inv(Wvw)[3:, 3:])
        Rf = np.ascontiguousarray(np.linalg.inv(Wvw)[:3, :3] / 0.05 ** 4)
        xvw_a = rng.normal(0, 1e-3, 6)
        xvw_b = xvw_a.copy()
        ga, ba, gb, bb = np.zeros(3), np.zeros(3), np.zeros(3), np.zeros(3)
        args = lambda X, P, ge, be, xv: (_p(a_s), _p(a_v), C.c_double(1.0), _p(Rot), _p(X), _p(P), _p(Qg), _p(Qrot), _p(Qbias),
                                         C.c_double(5e-6), C.c_double(1e10), _p(Rs), _p(Rf), _p(ge), _p(be), _p(Wvw), _p(xv),
                                         C.c_double(g))
        ka = L.rb_hostmath_est_ka_gmek_bias(*args(Xa, Pa, ga, ba, xvw_a))
        kb = R.ref_est_ka_gmek_bias(*args(Xb, Pb, gb, bb, xvw_b))
        assert np.isfinite(ka) and abs(ka - kb) <= 1e-8 * max(1.0, abs(kb)), (k, ka, kb)
        assert np.allclose(Xa, Xb, rtol=1e-8, atol=1e-10), (k, Xa, Xb)
        assert np.allclose(Pa, Pb, rtol=1e-6, atol=1e-18)
        assert np.allclose(xvw_a, xvw_b, rtol=1e-8, atol=1e-12)
        assert np.allclose(ga, gb, rtol=1e-9) and np.allclose(ba, bb, rtol=1e-7, atol=1e-12)

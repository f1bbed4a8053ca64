"""SURVEY.md 8(a) rows K6 / K13 (IMU mode, config 3): TryVel, Minimizer_V<double>, ExtRotVel against the unmodified
reference on a seeded frame pair.  Sums over keylines: rel 1e-10 (fixed-order reduction vs sequential adds); forward
matches and residual buffer: exact; minimiser output: 1e-9."""
import numpy as np
import pytest

from parity_util import DOG_THRESH, EUROC_CFG, PLANE_FIT, POS_NEG

pytestmark = pytest.mark.gpu


def test_tryvel_minimizer_v_extrotvel(built):
    from oracle import refapi
    from rebvo_b200 import capi, synth
    if not refapi.available():
        pytest.skip("oracle/_ref not built")
    cfg = EUROC_CFG
    cam = cfg["cam"]
    seq = synth.Sequence(w=cam["w"], h=cam["h"], seed=7, zf=cam["zfx"])
    f0, f1 = seq.frame(20)[1], seq.frame(21)[1]
    refs = [refapi.RefMap(cam["w"], cam["h"], cam["ppx"], cam["ppy"], cam["zfx"], cam["zfy"], cfg["sigma0"],
                          cfg["ksigma"]) for _ in range(2)]
    t, l = 0.012, 0
    for r, fr in zip(refs, (f0, f1)):
        r.rgb2bw(fr)
        r.build()
        _, t, l = r.detect(PLANE_FIT, POS_NEG, DOG_THRESH, cfg["kl_max"], t, l, cfg["kl_ref"], 0.0, 1, 0)
    old_r, new_r = refs
    _, rt_new = new_r.reestimate(cfg["track_points"], 100)
    _, rt_old = old_r.reestimate(cfg["track_points"], 100)
    rng = np.random.default_rng(2)
    kl = old_r.keylines()
    kl["rho"] = rng.uniform(0.7, 1.5, len(kl))
    kl["s_rho"] = rng.uniform(0.05, 0.5, len(kl))
    kl["m_num"] = rng.integers(0, 6, len(kl))
    old_r.set_keylines(kl)
    ctx = capi.Ctx(cam, cfg["sigma0"], cfg["ksigma"])
    old_g, new_g = ctx.new_map(), ctx.new_map()
    old_g.load_keylines(old_r.keylines(), old_r.mask())
    new_g.load_keylines(new_r.keylines(), new_r.mask())
    new_r.build_field(cfg["radius"], rt_new)
    new_g.build_field(cfg["radius"], rt_new)
    q = old_r.quantile(1e-3, 20.0, 0.9, 100)
    k0 = old_r.knum()
    res = np.zeros(k0)
    for V in (np.zeros(3), np.array([0.003, -0.002, 0.004]), np.array([-0.005, 0.001, -0.6])):
        s_r, J_r, F_r, res_r = new_r.try_vel(old_r, V, cfg["match_thresh"], q, 0, res, 2.0, rt_old)
        s_g, J_g, F_g, res_g = new_g.try_vel(old_g, V, cfg["match_thresh"], q, 0, res, 2.0, rt_old)
        assert np.isclose(s_r, s_g, rtol=1e-10), (s_r, s_g)
        assert np.allclose(J_r, J_g, rtol=1e-9, atol=1e-9 * np.abs(J_r).max())
        assert np.allclose(F_r, F_g, rtol=1e-9, atol=1e-9 * np.abs(F_r).max())
        assert np.array_equal(old_r.keylines()["m_id_f"], old_g.keylines()["m_id_f"])
        assert np.array_equal(res_r, res_g), "in-place residual buffer differs in %d entries" % int((res_r != res_g).sum())
        res = res_r
    m_r = new_r.minimizer_v(old_r, np.zeros(3), cfg["match_thresh"], 5, q, 0, 2.0, rt_old)
    m_g = new_g.minimizer_v(old_g, np.zeros(3), cfg["match_thresh"], 5, q, 0, 2.0, rt_old)
    print("Minimizer_V ref V", m_r["V"], "F", m_r["F"])
    assert np.allclose(m_r["V"], m_g["V"], rtol=1e-8, atol=1e-11)
    assert np.isclose(m_r["F"], m_g["F"], rtol=1e-9)
    assert np.allclose(m_r["RVel"], m_g["RVel"], rtol=1e-7, atol=1e-9 * np.abs(m_r["RVel"]).max())
    assert np.array_equal(old_r.keylines()["m_id_f"], old_g.keylines()["m_id_f"])
    # ExtRotVel on the forward matches
    old_r.forward_match(new_r)
    old_g.forward_match(new_g)
    ok_r, Wx_r, Rx_r, X_r = new_r.ext_rot_vel(m_r["V"], 1.0, 2.0)
    ok_g, Wx_g, Rx_g, X_g = new_g.ext_rot_vel(m_r["V"], 1.0, 2.0)
    assert ok_r and ok_g
    assert np.allclose(Wx_r, Wx_g, rtol=1e-9, atol=1e-9 * np.abs(Wx_r).max())
    assert np.allclose(X_r, X_g, rtol=1e-6, atol=1e-9 * np.abs(X_r).max())
    assert np.allclose(Rx_r, Rx_g, rtol=1e-6, atol=1e-9 * np.abs(Rx_r).max())
    old_g.close()
    new_g.close()
    ctx.close()

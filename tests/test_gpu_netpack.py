"""Wire egress: rb_map_pack_net_keylines against the reference's own packer
(copy_net_keyline + copy_net_keyline_nextid, src/CommLib/net_keypoint.cpp:29-107) on the same edge map -- byte for byte."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_net_keylines_bytes_equal_reference(built):
    from oracle import refapi
    from rebvo_b200 import capi, synth
    if not refapi.available():
        pytest.skip("compiled reference not available")
    cam = synth.EUROC
    ts, fr = synth.Sequence(w=cam["w"], h=cam["h"], seed=5, zf=cam["zfx"]).frames(6)
    # a real tracked map (depths, flow, match counts spread over their ranges): last map of a short pipeline run
    pl = capi.Pipeline(capi.default_params(cam), max_batch=6)
    pl.push(fr, ts)
    gmap = pl.map(0)
    kl, mask = gmap.keylines(), gmap.mask()
    assert len(kl) > 10000 and (kl["m_num"] > 0).sum() > 1000
    # extremes of the clamps: huge / tiny depths, long flows, large match counts
    rng = np.random.default_rng(1)
    kl = kl.copy()
    sel = rng.choice(len(kl), 400, replace=False)
    kl["rho"][sel[:100]] = rng.uniform(6.0, 20.0, 100)
    kl["rho"][sel[100:200]] = rng.uniform(5e-5, 1e-3, 100)
    kl["s_rho"][sel[200:300]] = 20.0
    kl["m_num"][sel[300:350]] = rng.integers(200, 400, 50)
    kl["p_m_0"][sel[350:], 0] += rng.uniform(-30, 30, 50).astype(np.float32)
    r = refapi.RefMap(cam["w"], cam["h"], cam["ppx"], cam["ppy"], cam["zfx"], cam["zfy"], 3.56359, 1.2599)
    r.set_keylines(kl)
    r.set_mask(mask, len(kl))     # (sets the reference object's keyline count)
    ctx = capi.Ctx(cam, 3.56359, 1.2599, kl_capacity=40000)
    g = ctx.new_map()
    g.load_keylines(kl, mask)
    for k_prof in (1.0, 0.37, 2.5):
        ref = r.pack_net(k_prof)
        got = g.pack_net(k_prof)
        assert ref.shape == got.shape == (len(kl), 15)
        bad = np.nonzero((ref != got).any(1))[0]
        assert len(bad) == 0, (k_prof, bad[:5], ref[bad[:3]], got[bad[:3]])
    # capacity below kn: truncated like copy_net_keyline's kl_size
    got = g.pack_net(1.0, capacity=5000)
    assert got.shape == (5000, 15)
    pl.close()

"""rb_map_clone: edge_finder / global_tracker copy constructors (edge_finder.cpp:42-52, global_tracker.cpp:42-47) as a
device-side deep copy: equal content, independent storage."""
import numpy as np
import pytest

from parity_util import DOG_THRESH, EUROC_CFG, PLANE_FIT, POS_NEG

pytestmark = pytest.mark.gpu


def test_clone_is_deep_and_equal(built):
    from rebvo_b200 import capi, synth
    cfg = EUROC_CFG
    cam = cfg["cam"]
    seq = synth.Sequence(w=cam["w"], h=cam["h"], seed=7, zf=cam["zfx"])
    ctx = capi.Ctx(cam, cfg["sigma0"], cfg["ksigma"])
    m = ctx.new_map()
    m.upload_rgb(seq.frame(3)[1])
    m.dog_build()
    det = capi.DetectParams(PLANE_FIT, POS_NEG, DOG_THRESH, cfg["kl_max"], cfg["kl_ref"], cfg["gain"], cfg["tmax"],
                            cfg["tmin"])
    kn, _, _ = m.detect(det, cfg["thresh"], 0)
    assert kn > 5000
    m.build_field(cfg["radius"], 0.0)
    c = m.clone()
    a_kl, a_mask, a_field = m.keylines(), m.mask(), m.field()
    b_kl, b_mask, b_field = c.keylines(), c.mask(), c.field()
    assert len(b_kl) == kn and np.array_equal(a_mask, b_mask) and np.array_equal(a_field, b_field)
    for f in a_kl.dtype.names:
        assert np.array_equal(a_kl[f], b_kl[f]), f
    # independent storage: rotating the original leaves the copy untouched
    R = np.array([[0.9998, -0.02, 0.0], [0.02, 0.9998, 0.0], [0.0, 0.0, 1.0]])
    m.rotate(R)
    assert not np.array_equal(m.keylines()["p_m"], a_kl["p_m"])
    assert np.array_equal(c.keylines()["p_m"], a_kl["p_m"])
    c.close()
    m.close()
    ctx.close()

"""CPU suite: pins the oracle.  (1) The CPU restatement oracle/rebvo_oracle.cpp against the golden vectors that
tests/golden/make_golden.py generated from the unmodified reference; (2) the compiled reference itself against
the same vectors when oracle/_ref is present; (3) restatement vs compiled reference on a second, larger seeded
input.  Integer / float32 / per-keyline float64 results are compared bit for bit; only what passes through the
6x6 SVD solve (LAPACK in the reference) gets a 1e-9 tolerance."""
import os

import numpy as np
import pytest

from flow import SMALL, compare, run_flow, small_frames

GOLD = os.path.join(os.path.dirname(__file__), "golden", "flow_small.npz")
TOL = ("min_V", "min_W", "min_RVel", "min_RW0", "min_W_X", "min_scalars")


@pytest.fixture(scope="module")
def golden():
    z = np.load(GOLD)
    return {k: z[k] for k in z.files}


def _override(g):
    return dict(V=g["min_V"], W=g["min_W"], RVel=g["min_RVel"], RW0=g["min_RW0"])


def test_golden_inputs_reproducible(golden):
    f0, f1 = small_frames()
    assert np.array_equal(f0, golden["f0"]) and np.array_equal(f1, golden["f1"])


def test_port_matches_golden(golden):
    from oracle import portapi
    out = run_flow(portapi.PortMap, SMALL, golden["f0"], golden["f1"], portapi.so3_exp, _override(golden))
    ref = {k: v for k, v in golden.items() if k not in ("f0", "f1")}
    fails = compare(ref, out, tol_keys=TOL)
    assert not fails, "\n".join(fails)
    # the minimiser itself: same LM path, different 6x6 solver
    assert np.allclose(out["min_V"], golden["min_V"], rtol=1e-9, atol=1e-12)
    assert np.allclose(out["min_W"], golden["min_W"], rtol=1e-9, atol=1e-12)


def test_reference_matches_golden(golden):
    from oracle import refapi
    if not refapi.available():
        pytest.skip("oracle/_ref not built")
    out = run_flow(refapi.RefMap, SMALL, golden["f0"], golden["f1"], refapi.so3_exp, _override(golden))
    ref = {k: v for k, v in golden.items() if k not in ("f0", "f1")}
    fails = compare(ref, out, tol_keys=TOL)
    assert not fails, "\n".join(fails)


def test_port_matches_reference_qvga():
    from oracle import portapi, refapi
    from rebvo_b200 import synth
    if not refapi.available():
        pytest.skip("oracle/_ref not built")
    cfg = dict(SMALL, cam=dict(w=320, h=240, zfx=260.0, zfy=258.0, ppx=161.0, ppy=118.5), kl_max=9000, kl_ref=5000,
               track_points=4000, radius=20, sigma0=3.56359)
    f0, f1 = synth.frame_pair(seed=23, w=320, h=240, nrect=90, shift=(-1.6, 0.9))
    a = run_flow(refapi.RefMap, cfg, f0, f1, refapi.so3_exp)
    ov = dict(V=a["min_V"], W=a["min_W"], RVel=a["min_RVel"], RW0=a["min_RW0"])
    b = run_flow(portapi.PortMap, cfg, f0, f1, portapi.so3_exp, ov)
    fails = compare(a, b, tol_keys=TOL)
    assert not fails, "\n".join(fails)
    assert a["f0_kl"].shape[0] > 2000


def test_kl_max_truncation_and_empty_image():
    """Edge cases of build_mask: the kl_max cut clears the rest of the mask; a flat image yields no keylines."""
    from oracle import portapi
    cam = SMALL["cam"]
    f0, _ = small_frames()
    m = portapi.PortMap(cam["w"], cam["h"], cam["ppx"], cam["ppy"], cam["zfx"], cam["zfy"], 1.7818, 1.2599)
    m.rgb2bw(f0)
    m.build()
    kn_full, _, _ = m.detect(2, 0.4, 0.0952598689, 3000, 0.012, 0, 1500, 0.0, 1, 0)
    full_mask = m.mask().copy()
    kn, _, _ = m.detect(2, 0.4, 0.0952598689, 500, 0.012, 0, 1500, 0.0, 1, 0)
    assert kn == 500 and kn_full > 500
    mask = m.mask()
    assert mask.max() == 499 and (mask >= 0).sum() == 500
    assert np.array_equal(mask[(full_mask >= 0) & (full_mask < 500)], full_mask[(full_mask >= 0) & (full_mask < 500)])
    flat = np.full((cam["h"], cam["w"], 3), 90, np.uint8)
    m.rgb2bw(flat)
    m.build()
    kn, _, _ = m.detect(2, 0.4, 0.0952598689, 3000, 0.012, 0, 1500, 0.0, 1, 0)
    assert kn == 0 and (m.mask() == -1).all()
    assert np.abs(m.plane("dog")).max() < 1e-3


def test_box_plan_known_answers():
    """Kovesi box widths / achieved sigmas of iigauss::iigauss."""
    from oracle import portapi
    for sigma0, want in ((1.7818, ([3, 3, 5], [3, 5, 5], 1.825742, 2.160247)),
                         (3.56359, ([7, 7, 7], [9, 9, 9], 3.464102, 4.472136))):
        m = portapi.PortMap(64, 64, 32, 32, 50, 50, sigma0, 1.2599)
        d, s = m.box_plan()
        assert d[0].tolist() == want[0] and d[1].tolist() == want[1]
        assert abs(s[0] - want[2]) < 1e-6 and abs(s[1] - want[3]) < 1e-6

"""image_undistort::undistort<true> on RGB24 -- integer-exact, so the CUDA remap must equal
the reference's bit for bit, including the frame border where taps fall outside the image."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# EuRoC cam0 radial-tangential coefficients (Kc2, Kc4, Kc6, P1, P2), app/rebvorun/GlobalConfig_EuRoC_2.txt
EUROC_KC = [-0.28340811, 0.07395907, 0.0, 0.00019359, 1.76187114e-05]


@pytest.mark.parametrize("kc", [EUROC_KC, [0.12, -0.05, 0.01, -0.002, 0.003], [0, 0, 0, 0, 0]])
def test_undistort_matches_reference(built, kc):
    from oracle import refapi
    from rebvo_b200 import capi, synth
    if not refapi.available():
        pytest.skip("oracle/_ref not built")
    cam = synth.EUROC
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (cam["h"], cam["w"], 3), dtype=np.uint8)
    img2 = synth.Sequence(w=cam["w"], h=cam["h"], seed=3).frame(2)[1]
    ctx = capi.Ctx(cam, 3.56359, 1.2599)
    und = capi.Undistort(ctx, kc)
    for im in (img, img2):
        ref = refapi.undistort_rgb(cam, kc, im)
        got = und.apply(im)
        assert np.array_equal(ref, got), "mismatch in %d bytes" % int((ref != got).sum())
    und.close()
    ctx.close()

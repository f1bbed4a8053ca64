"""Config 3 (IMU fusion, ImuMode=2: gyro pre-rotation, Minimizer_V, ExtRotVel, BiasCorrect, scale / gravity / bias
filter) end to end.  The reference's own SecondThread IMU branch (rebvo_second_t.cpp:182-336, unmodified, with its
ScaleEstimator and ImuGrabber) runs once on the reference's CPU hot path (oracle/_ref/ref_rebvo) and once on
librebvo_b200 through the shim (oracle/_ref/shim_rebvo): same frames, same synthetic IMU csv."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
NF = 50


def test_imu_mode_trajectory(built, tmp_path):
    from oracle import refapi
    from rebvo_b200 import synth
    shim_exe = os.path.join(os.path.dirname(refapi.EXE), "shim_rebvo")
    if not (os.path.exists(shim_exe) and os.path.exists(refapi.EXE)):
        pytest.skip("oracle/_ref/shim_rebvo / ref_rebvo not built (need the reference sources)")
    cam = synth.EUROC
    seq = synth.Sequence(w=cam["w"], h=cam["h"], seed=7, zf=cam["zfx"])
    ts, fr = seq.frames(NF)
    path = str(tmp_path / "frames.bin")
    synth.write_frames_file(path, ts, fr)
    csv = str(tmp_path / "imu.csv")
    synth.write_imu_csv(csv, synth.imu_samples(seq, NF))
    kv = dict(ImuMode=2, ImuFile=csv, ImuTimeScale=1, InitBias=1, InitBiasFrameNum=5)
    _, ref = refapi.run_full_rebvo(path, path + ".ref", kv)
    _, rec = refapi.run_full_rebvo(path, path + ".shim", kv, exe=shim_exe)
    n = min(len(ref), len(rec))
    assert n >= NF - 2
    assert np.array_equal(ref["kn"][:n], rec["kn"][:n])
    dm = np.abs(ref["matches"][1:n].astype(int) - rec["matches"][1:n].astype(int))
    e = np.sqrt(((ref["Pos"][:n] - rec["Pos"][:n]) ** 2).sum(1))
    print("IMU mode: max |d matches| %d, ATE %.3e m, max %.3e m, K ref %s gpu %s" %
          (dm.max(), np.sqrt((e ** 2).mean()), e.max(), ref["K"][n - 1], rec["K"][n - 1]))
    assert np.isfinite(ref["Pos"][:n]).all() and np.isfinite(rec["Pos"][:n]).all()
    assert dm.max() <= 0.005 * ref["matches"][1:n].max()
    assert e.max() <= 1e-3                        # bar of north_star
    assert np.allclose(ref["K"][:n], rec["K"][:n], rtol=1e-3, atol=1e-9)

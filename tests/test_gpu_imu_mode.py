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
    assert np.allclose(ref["K"][:n], rec["K"][:n], rtol=1e-6, atol=1e-12)


def test_imu_mode_pipeline_vs_reference(built, tmp_path):
    """The product's own IMU-mode flow (rb_pipeline_set_imu: csrc/imu_flow.cuh + imu_filter.h) against the reference's CPU
    build with ImuMode=2 on the same frames and the same synthetic IMU samples."""
    from oracle import refapi
    from rebvo_b200 import capi, synth
    if not os.path.exists(refapi.EXE):
        pytest.skip("oracle/_ref/ref_rebvo not built")
    cam = synth.EUROC
    seq = synth.Sequence(w=cam["w"], h=cam["h"], seed=7, zf=cam["zfx"])
    ts, fr = seq.frames(NF)
    smp = synth.imu_samples(seq, NF)
    path = str(tmp_path / "frames.bin")
    synth.write_frames_file(path, ts, fr)
    csv = str(tmp_path / "imu.csv")
    synth.write_imu_csv(csv, smp)
    p = capi.default_params(cam)
    kv = refapi.ref_params_from(p, ImuMode=2, ImuFile=csv, ImuTimeScale=1, InitBias=1, InitBiasFrameNum=5)
    _, ref = refapi.run_full_rebvo(path, path + ".ref", kv)
    pl = capi.Pipeline(p, max_batch=10)
    pl.set_imu(smp, capi.default_imu_params(InitBias=1, InitBiasFrameNum=5))
    nav = np.concatenate([pl.push(fr[s:s + 10], ts[s:s + 10]) for s in range(0, NF, 10)])
    pl.close()
    n = min(len(ref), len(nav))
    assert n >= NF - 2
    assert np.array_equal(ref["kn"][:n], nav["kn"][:n])
    dm = np.abs(ref["matches"][1:n].astype(int) - nav["matches"][1:n].astype(int))
    e = np.sqrt(((ref["Pos"][:n] - nav["Pos"][:n]) ** 2).sum(1))
    print("IMU pipeline: max |d matches| %d, ATE %.3e m, max %.3e m, K ref %.12g gpu %.12g" %
          (dm.max(), np.sqrt((e ** 2).mean()), e.max(), ref["K"][n - 1], nav["K"][n - 1]))
    assert np.isfinite(nav["Pos"][:n]).all()
    assert dm.max() <= 0.005 * ref["matches"][1:n].max()
    assert e.max() <= 1e-3
    assert np.allclose(ref["K"][1:n], nav["K"][1:n], rtol=1e-6, atol=1e-12)   # (record 0 is the untracked first frame)
    assert np.abs(ref["PoseLie"][:n] - nav["PoseLie"][:n]).max() <= 1e-3

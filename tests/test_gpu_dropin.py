"""The drop-in boundary, executed: oracle/_ref/shim_rebvo is the reference's own REBVO class -- rebvo.cpp, rebvo_first_t.cpp,
rebvo_second_t.cpp, rebvo_third_t.cpp, keyframe / kfvo / pose_graph, CommLib, VideoLib, all UNMODIFIED and compiled where
they lie -- built against include/rebvo_b200_shim.hpp instead of the six hot-path translation units (oracle/build_ref.py
build_shim_rebvo).  Its three pipeline threads call the stage-level C ABI concurrently (detector thread on slot i, tracker
thread on slots i-1 / i-2).  It must reproduce the reference's CPU build, run after run."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
NF = 60


def test_unmodified_three_thread_rebvo_on_the_gpu_library(built, tmp_path):
    from oracle import refapi
    from rebvo_b200 import capi, synth
    shim_exe = os.path.join(os.path.dirname(refapi.EXE), "shim_rebvo")
    if not (os.path.exists(shim_exe) and os.path.exists(refapi.EXE)):
        pytest.skip("oracle/_ref/shim_rebvo / ref_rebvo not built (need the reference sources)")
    cam = synth.EUROC
    ts, fr = synth.Sequence(w=cam["w"], h=cam["h"], seed=13, zf=cam["zfx"]).frames(NF)
    path = str(tmp_path / "frames.bin")
    synth.write_frames_file(path, ts, fr)
    _, ref = refapi.run_full_rebvo(path, path + ".ref")
    runs = []
    for k in range(3):
        info, rec = refapi.run_full_rebvo(path, path + ".shim%d" % k, exe=shim_exe)
        print("run %d: %d callbacks, %.1f frames/s (three host threads, stage-level calls)" % (k, len(rec), info["fps"]))
        runs.append(rec.copy())
    n = min(len(ref), min(len(r) for r in runs))
    assert n >= NF - 2
    for k in range(1, 3):   # thread timing must not change the result
        for f in ("kn", "matches", "Pos", "PoseLie", "Kp", "est_ok"):
            assert np.array_equal(runs[0][f][:n], runs[k][f][:n]), (k, f)
    rec = runs[0]
    # same arithmetic as the device-resident pipeline (stage by stage instead of fused): same poses
    pl = capi.Pipeline(capi.default_params(cam), max_batch=20)
    nav = np.concatenate([pl.push(fr[s:s + 20], ts[s:s + 20]) for s in range(0, NF, 20)])[:n]
    pl.close()
    assert np.array_equal(nav["kn"], rec["kn"][:n]) and np.array_equal(nav["matches"][1:], rec["matches"][1:n])
    assert np.abs(nav["Pos"] - rec["Pos"][:n]).max() <= 1e-9
    # against the CPU reference: the detector is bit-exact; the tracker's sums are added in another order and its 6x6
    # solves / sin / cos differ in the last bit, so a frame whose LM decision sits on a knife edge (which init try wins,
    # accept / reject, a pixel rounding) can pick the other branch -- this stream has one such frame (30) -- and the two
    # trajectories then differ by far less than the bar of north_star (ATE <= 1e-3 m), not by rounding only.
    assert np.array_equal(ref["kn"][:n], rec["kn"][:n])
    dm = np.abs(ref["matches"][1:n].astype(int) - rec["matches"][1:n].astype(int))
    assert dm.max() <= 0.005 * ref["matches"][1:n].max(), dm.max()
    e = np.sqrt(((ref["Pos"][:n] - rec["Pos"][:n]) ** 2).sum(1))
    ate = float(np.sqrt((e ** 2).mean()))
    first = int(np.nonzero(e > 1e-9)[0][0]) if (e > 1e-9).any() else -1
    print("unmodified REBVO on librebvo_b200 vs CPU reference: ATE %.3e m over %d frames (first frame above 1e-9 m: %d)" % (ate, n, first))
    assert ate <= 1e-3 and e.max() <= 1e-3
    assert np.abs(ref["PoseLie"][:n] - rec["PoseLie"][:n]).max() <= 1e-3

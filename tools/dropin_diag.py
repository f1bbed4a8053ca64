import os, sys, subprocess
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refapi
from rebvo_b200 import synth, capi
cam = synth.EUROC
NF = 60
ts, fr = synth.Sequence(w=cam["w"], h=cam["h"], seed=13, zf=cam["zfx"]).frames(NF)
path = "/tmp/dd_frames.bin"
synth.write_frames_file(path, ts, fr)
_, ref = refapi.run_full_rebvo(path, path + ".ref")
_, shm = refapi.run_full_rebvo(path, path + ".shim", exe=os.path.join(os.path.dirname(refapi.EXE), "shim_rebvo"))
exe = os.path.join(os.path.dirname(refapi.EXE), "shim_driver")
r = subprocess.run([exe, path, path + ".drv"], capture_output=True, text=True, timeout=600)
with open(path + ".drv", "rb") as f:
    n, sz = np.frombuffer(f.read(8), np.int32)
    drv = np.frombuffer(f.read(), refapi.OUTREC, count=n)
pl = capi.Pipeline(capi.default_params(cam), max_batch=20)
nav = np.concatenate([pl.push(fr[s:s + 20], ts[s:s + 20]) for s in range(0, NF, 20)])
n = min(len(ref), len(shm), len(drv), len(nav))
print("n", n)
for name, a in (("shim_rebvo", shm), ("shim_driver", drv)):
    mm = np.nonzero(ref["matches"][1:n] != a["matches"][1:n])[0]
    print(name, "first matches mismatch at", (mm[0] + 1) if len(mm) else None, "kn equal", np.array_equal(ref["kn"][:n], a["kn"][:n]))
    d = np.sqrt(((ref["Pos"][:n] - a["Pos"][:n]) ** 2).sum(1))
    print("   pos err per frame (first 12 nonzero):", [(int(i), float("%.2e" % d[i])) for i in np.nonzero(d > 1e-12)[0][:12]])
    print("   Kp diff first:", [(int(i)) for i in np.nonzero(ref["Kp"][:n] != a["Kp"][:n])[0][:8]])
mm = np.nonzero(ref["matches"][1:n] != nav["matches"][1:n])[0]
print("pipeline first mismatch", mm[:3])
i = (np.nonzero(ref["matches"][1:n] != shm["matches"][1:n])[0][0] + 1) if np.any(ref["matches"][1:n] != shm["matches"][1:n]) else 1
for j in range(max(1, i - 2), min(n, i + 3)):
    print(j, "ref", ref["matches"][j], ref["Kp"][j], ref["est_ok"][j], ref["s_rho_p"][j], "| shim", shm["matches"][j], shm["Kp"][j], shm["est_ok"][j], shm["s_rho_p"][j])

import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rebvo_b200 import capi, synth
capi.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_prof', 'librebvo_b200_dbg.so')
cam = synth.EUROC
B = 16
seq = synth.Sequence(w=cam["w"], h=cam["h"], seed=7, zf=cam["zfx"])
ts, fr = seq.frames(B * 4)
pl = capi.Pipeline(capi.default_params(cam), max_batch=B)
L = capi.lib()
buf = np.zeros(8192, np.uint64); n = C.c_uint(0)
for s in range(4):
    nav = pl.push(fr[s * B:(s + 1) * B], ts[s * B:(s + 1) * B])
    L.rb_debug_fetch_trace(buf.ctypes.data_as(C.c_void_p), C.byref(n))
print('events in last push', n.value)
ev = [(int(v >> np.uint64(56)), int(v & np.uint64(0x00FFFFFFFFFFFFFF))) for v in buf[:n.value]]
ev.sort(key=lambda x: x[1])
t0 = ev[0][1]
names = {8: 'rega_end', 9: 'ekf_end', 11: 'rot_end', 1: 'trk_begin', 2: 'min_begin', 3: 'min_end', 4: 'match_end', 5: 'trk_end', 6: 'det_begin', 7: 'det_end'}
# per-tag sequences
byt = {k: [t - t0 for tag, t in ev if tag == k] for k in names}
def us(x): return x / 1e3
for i in range(min(B, len(byt[5]))):
    row = []
    for k in (6, 7, 1, 2, 3, 4, 5):
        row.append('%s=%.1f' % (names[k], us(byt[k][i])) if i < len(byt[k]) else '')
    print(i, ' '.join(row))
d = np.diff(byt[5]); print('frame period us: mean %.1f' % us(d.mean()))
print('means us: pre->min_begin %.1f  minimiser %.1f  min_end->match_end %.1f  match_end->trk_end %.1f  detect %.1f  trk_end->next trk_begin %.1f' % (
    us(np.mean(np.array(byt[2]) - np.array(byt[1]))), us(np.mean(np.array(byt[3]) - np.array(byt[2]))),
    us(np.mean(np.array(byt[4]) - np.array(byt[3]))), us(np.mean(np.array(byt[5]) - np.array(byt[4]))),
    us(np.mean(np.array(byt[7]) - np.array(byt[6]))), us(np.mean(np.array(byt[1][1:]) - np.array(byt[5][:-1])))))

A = lambda k: np.array(byt[k][:len(byt[5])])
print('fm+rot %.1f dmatch %.1f | reg+ekf %.1f rescale+finish %.1f' % (us(np.mean(A(11) - A(3))), us(np.mean(A(4) - A(11))), us(np.mean(A(9) - A(4))), us(np.mean(A(5) - A(9)))))

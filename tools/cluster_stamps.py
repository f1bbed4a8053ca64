"""Phase timings of the cluster minimiser (profiling build, tools/build_prof.py): clock64 stamps per CTA and round."""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["REBVO_B200_NO_GRAPH"] = "1"
from rebvo_b200 import capi, synth
capi.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_prof', 'librebvo_b200_dbg.so')
cam = synth.EUROC
seq = synth.Sequence(w=cam["w"], h=cam["h"], seed=7, zf=cam["zfx"])
ts, fr = seq.frames(8)
pl = capi.Pipeline(capi.default_params(cam), max_batch=8)
nav = pl.push(fr, ts)
print('kn', nav['kn'], 'pos', nav['Pos'][-1])
dbg = np.zeros(256 * 16, np.int64)
L = capi.lib()
print('fetch rc', L.rb_debug_fetch(dbg.ctypes.data_as(C.c_void_p)))
d = dbg.reshape(32, 16, 8)
names = ['exp', 'eval', 'xchg', 'totals', 'lm']
for b in (0, 7, 15):
    k = d[b, 15]
    print('CTA %d: entry->pdl %d, prologue %d, rounds %d, epilogue-sync %d, total %d cycles' %
          (b, k[1] - k[0], k[2] - k[1], k[3] - k[2], k[4] - k[3], k[4] - k[0]))
    for e in range(12):
        r = d[b, e]
        if r[0] == 0:
            continue
        nxt = d[b, e + 1][0] if e + 1 < 12 and d[b, e + 1][0] else k[3]
        print('   round %2d: ' % e + ' '.join('%s=%d' % (names[i], r[i + 1] - r[i]) for i in range(5)) +
              ' | round %d' % (nxt - r[0]))

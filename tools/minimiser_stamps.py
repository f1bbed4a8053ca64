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
d = dbg.reshape(256, 16)
print('n_act', int(d[255, 15]))
names = ['exp', 'body', 'tail+ll', 'gather', 'carries', 'reduce', 'lm_step', 'epilogue']
for b in (0, 1, 30, 58):
    row = d[b]
    print('block', b, ' '.join('%s=%d' % (names[k], row[k + 1] - row[k]) for k in range(0, 8) if row[k + 1] and row[k] and row[k + 1] > row[k]))
na = int(d[255, 15])
pub = int(d[0][12]); obs = np.array([int(d[b][12]) for b in range(1, na)]); sto = np.array([int(d[b][13]) for b in range(na)]); gath = int(d[0][14])
print('globaltimer ns: request published -> observed by workers: min %d mean %d max %d' % ((obs - pub).min(), (obs - pub).mean(), (obs - pub).max()))
print('worker observed -> sums stored: mean %d ; last store -> master gathered %d ; first store -> gathered %d' % ((sto[1:] - obs).mean(), gath - sto.max(), gath - sto.min()))
print('master publish -> gathered (one evaluation minus LM/exp): %d ns' % (gath - pub))
print('master last eval total', d[0][7] - d[0][0])

"""Phase timings of the cluster minimiser (profiling build, tools/build_prof.py): clock64 stamps of thread 0 per CTA and round."""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["REBVO_B200_NO_GRAPH"] = "1"
from rebvo_b200 import capi, synth
capi.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_prof' + (sys.argv[1] if len(sys.argv) > 1 else ''), 'librebvo_b200_dbg.so')
cam = synth.EUROC
seq = synth.Sequence(w=cam["w"], h=cam["h"], seed=7, zf=cam["zfx"])
ts, fr = seq.frames(8)
pl = capi.Pipeline(capi.default_params(cam), max_batch=8)
nav = pl.push(fr, ts)
print('kn', nav['kn'], 'pos', nav['Pos'][-1])
dbg = np.zeros(256 * 16, np.int64)
L = capi.lib()
print('fetch rc', L.rb_debug_fetch(dbg.ctypes.data_as(C.c_void_p)))
d = dbg.reshape(16, 16, 16)
names = ['eval', 'send', 'wait', 'totals', 'ingest', 'lm', 'expR']
for b in (0, 15):
    k = d[b, 15]
    print('CTA %d: entry->pdl %d, prologue %d, rounds %d, epilogue-sync %d, total %d cycles' %
          (b, k[1] - k[0], k[2] - k[1], k[3] - k[2], k[4] - k[3], k[4] - k[0]))
    for e in range(12):
        r = d[b, e]
        if r[0] == 0:
            continue
        nxt = d[b, e + 1][0] if e + 1 < 12 and d[b, e + 1][0] else k[3]
        print('   round %2d: ' % e + ' '.join('%s=%d' % (names[i], r[i + 1] - r[i]) for i in range(7) if r[i + 1] and r[i]) +
              ' | round %d' % (nxt - r[0]))
b = d[0, 14]
if b[0]:
    bn = ['SE3', '1/pz', 'pixel', 'field', 'pack', 'residual', 'sqrt', '1/q', 'products']
    print('body stages (thread 0, last keyline iteration of the last evaluation): ' + ' '.join('%s=%d' % (bn[i], b[i + 1] - b[i]) for i in range(9) if b[i + 1] and b[i]))
if b[10]:
    print('eval (thread 0, last evaluation): first iteration %d, remaining iterations %d, wait for the CTA %d, butterfly %d, sync %d' %
          (b[11] - b[10], b[12] - b[11], b[13] - b[12], b[14] - b[13], b[15] - b[14]))
try:
    gt = np.zeros(128 * 16 * 2, np.int64)
    L.rb_debug_fetch_gt(gt.ctypes.data_as(C.c_void_p))
    g = gt.reshape(128, 16, 2)
    n = int((g[:, 5, 0] != 0).sum())
    print('inter-cluster exchange, %d CTAs (ns, %%globaltimer):' % n)
    for e in (3, 5, 7):
        pub, got = g[:n, e, 0], g[:n, e, 1]
        G = n // 16
        lat = []
        for b in range(n):
            r = b % 16
            partners = [c * 16 + r for c in range(G)]
            lat.append(got[b] - max(pub[q] for q in partners))
        print('  round %d: publish spread %d ns (first %d .. last +%d); gathered - last partner publish: min %d mean %d max %d ns; per cluster mean publish offset %s' %
              (e, pub.max() - pub.min(), 0, pub.max() - pub.min(), min(lat), int(np.mean(lat)), max(lat),
               [int(pub[c * 16:(c + 1) * 16].mean() - pub.min()) for c in range(G)]))
except Exception as ex:
    print('no gt stamps', ex)

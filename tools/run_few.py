"""Push a few frames through the pipeline (eager launches) -- target of ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["REBVO_B200_NO_GRAPH"] = "1"
from rebvo_b200 import capi, synth
cam = synth.EUROC
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
seq = synth.Sequence(w=cam["w"], h=cam["h"], seed=7, zf=cam["zfx"])
ts, fr = seq.frames(n)
pl = capi.Pipeline(capi.default_params(cam), max_batch=n)
nav = pl.push(fr, ts)
print('kn', nav['kn'][-1], 'pos', nav['Pos'][-1])

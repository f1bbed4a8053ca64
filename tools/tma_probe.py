import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rebvo_b200 import capi, synth
cam = synth.EUROC
rng = np.random.default_rng(0)
img = synth.to_rgb_u8(synth.rect_canvas(rng, cam["w"], cam["h"], 60), rng)
ctx = capi.Ctx(cam, 3.56359, 1.2599, kl_capacity=20000)
m = ctx.new_map()
m.upload_rgb(img)
m.dog_build()
d = m.plane("dog")
print("dog", float(np.abs(d).sum()))

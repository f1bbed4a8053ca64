"""Build an alternative librebvo_b200 with extra -D flags into rebvo_b200/alt/<tag>/ (A/B measurements on the GPU box)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rebvo_b200 import build as B
extra = [a for a in sys.argv[1:] if a.startswith('-D')]
tag = ''.join(a for a in sys.argv[1:] if not a.startswith('-'))
out = os.path.join(os.path.dirname(os.path.abspath(B.__file__)), 'alt', tag)
os.makedirs(out, exist_ok=True)
objs = []
for s in B.SOURCES:
    o = os.path.join(out, s.replace('.cu', '.o'))
    subprocess.check_call([B.NVCC] + B.FLAGS + extra + ['-c', os.path.join(B.CSRC, s), '-o', o])
    objs.append(o)
subprocess.check_call([B.NVCC, '-shared', '-o', os.path.join(out, 'librebvo_b200.so')] + objs + ['-lcudart'])
for o in objs:
    os.remove(o)
print('ok', out)

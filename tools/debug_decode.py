import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import qoaudio_b200 as q
from oracle import binding as ob
eng = q.default_engine()
data = open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "julien_baker_sprained_ankle.qoa"), "rb").read()
infos, outs = eng.decode_batch([data], 0)
_, ref, _ = ob.decode_iter(data)
got = outs[0]
bad = np.nonzero(got != ref)[0]
print("n", len(ref), "mismatches", len(bad))
if len(bad):
    pairs = bad // 2
    frames = pairs // 5120
    print("first bad sample idx", bad[:10], "frame", frames[:10], "in-frame sample", (pairs % 5120)[:10])
    uf, cnt = np.unique(frames, return_counts=True)
    print("bad frames (first 40):", list(zip(uf[:40], cnt[:40])))
    f0 = uf[0]
    insl = np.unique((pairs[frames == f0] % 5120) // 20)
    print("bad slices in frame", f0, ":", insl[:60])

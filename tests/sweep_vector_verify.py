"""Long random sweep (not collected by pytest): the C oracle's sorted + pruned search against the vectorised verifier on
random shapes (1-8 channels, music / noise / clipped, lengths around every frame and slice boundary).
  python tests/sweep_vector_verify.py [seconds]      # 420 s: 5268 cases, 10.5 M slices, no mismatch (round 2)
"""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch, signals
from oracle import binding as ob, vector_verify as vv
rng = np.random.default_rng(12345)
t0 = time.time(); cases = slices = 0
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 420.0
while time.time() - t0 < budget:
    ch = int(rng.integers(1, 9)); kind = ['music', 'noise'][int(rng.integers(0, 2))]
    n = int(rng.choice([1, 19, 20, 21, 5119, 5120, 5121, 10259, int(rng.integers(1, 30000))]))
    seed = int(rng.integers(0, 1 << 30)); S = int(rng.integers(1, 4))
    src = [signals.synth_pcm(seed, s, ch, n, kind) for s in range(S)]
    if rng.random() < 0.2:   # scale to extremes
        src = [np.clip(p.astype(np.int64) * 8, -32768, 32767).astype(np.int16) for p in src]
    enc = [ob.encode_all(p, ch, 44100, n)[1] for p in src]
    dec = [ob.decode_all(e)[1] for e in enc]
    q = torch.from_numpy(np.stack([np.frombuffer(e, dtype=np.uint8) for e in enc]).copy())
    r = vv.verify_encode(q, torch.from_numpy(np.stack(src).copy()), torch.from_numpy(np.stack(dec).copy()), ch, n, 44100)
    bad = {k: v for k, v in r.items() if k.startswith('bad') or k.startswith('chain')}
    if any(bad.values()) or r['unverifiable_slices']:
        print('MISMATCH', ch, n, kind, seed, S, r); 
    cases += 1; slices += r['slices']
print('cases', cases, 'slices', slices, 'clean')

"""The bodies of tests/test_gpu_verify.py, run on the CPU against a stand-in engine made of the oracle.

Those GPU tests drive full-size batches through the vectorised verifier; this dry run executes the very same test
functions (source loaded from the file, `device="cuda"` -> `"cpu"`, torch.cuda.* stubbed, tiny batch sizes because the
stub reports little free memory) so that everything but the CUDA engine itself — shapes, offsets, chunking, the
verifier's use — is exercised in the `-m "not gpu"` run as well.  The stand-in implements only the engine calls those
tests make; it lives here, in tests/, and is built from oracle/ (test infrastructure), never from the product."""
import ctypes
import os
import types

import numpy as np
import torch

import signals
from oracle import binding as ob
from test_vector_verify import random_payload_stream

HERE = os.path.dirname(os.path.abspath(__file__))


def _wr(ptr, b):
    ctypes.memmove(ptr, bytes(b), len(b))


def _rd(ptr, n):
    return ctypes.string_at(ptr, n)


class _Plan:
    def __init__(self, ptr, in_off, lens):
        self.dec = [ob.decode_all(_rd(ptr + int(o), int(l))) for o, l in zip(in_off, lens)]
        self.info_array = np.zeros(len(self.dec), dtype=[("samples", np.uint64)])
        self.info_array["samples"] = [d[0]["samples"] for d in self.dec]

    def run_dev(self, ptr, out_off, caps, stream=0):
        for (_, pcm), o in zip(self.dec, out_off):
            _wr(ptr + 2 * int(o), pcm.tobytes())
        return np.zeros(len(self.dec), dtype=np.int32)

    def close(self):
        pass


class _StandInEngine:
    def gen_qoa_dev(self, ptr, stride, n_streams, first, ch, rate, n, seed, stream=0):
        rng = np.random.default_rng(seed)
        for s in range(n_streams):
            _wr(ptr + s * stride, random_payload_stream(rng, ch, rate, n))

    def gen_pcm_dev(self, ptr, n_streams, first, ch, n, start, seed, kind="music", stream=0):
        for s in range(n_streams):
            _wr(ptr + 2 * s * n * ch, signals.synth_pcm(seed, first + s, ch, n, kind, start).tobytes())

    def encode_all_batch_dev(self, descs, d_pcm, pcm_off, pcm_len, d_out, out_off, out_cap, stream=0):
        st, ln = [], []
        for (c, r, n), po, pl, oo in zip(descs, pcm_off, pcm_len, out_off):
            pcm = np.frombuffer(_rd(d_pcm + 2 * int(po), 2 * int(pl)), dtype=np.int16)
            s, enc = ob.encode_all(pcm, c, r, n)
            _wr(d_out + int(oo), enc)
            st.append(s)
            ln.append(len(enc))
        return st, ln

    def decode_plan_dev(self, ptr, in_off, lens, mode, stream=0):
        return _Plan(ptr, in_off, lens)

    def decode_batch(self, streams, mode):
        out = [ob.decode_all(s) for s in streams]
        return [o[0] for o in out], [o[1] for o in out]


def test_gpu_verify_bodies_on_a_stand_in_engine(fixture_bytes, monkeypatch):
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)
    monkeypatch.setattr(torch.cuda, "mem_get_info", lambda: (int(1e9), int(1e9)))   # -> the smallest batch tier
    monkeypatch.setattr(torch.cuda, "current_stream", lambda: types.SimpleNamespace(cuda_stream=0))
    monkeypatch.setattr(torch.cuda, "synchronize", lambda: None)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self: self)
    src = open(os.path.join(HERE, "test_gpu_verify.py")).read()
    assert src.count('device="cuda"') >= 4
    mod = types.ModuleType("gpu_verify_dryrun")
    exec(compile(src.replace('device="cuda"', 'device="cpu"'), "test_gpu_verify.py", "exec"), mod.__dict__)
    eng, native = _StandInEngine(), types.SimpleNamespace(DECODE_ALL=1)
    mod.test_fixture_gpu_decode_matches_the_encoder_state_in_its_headers(eng, fixture_bytes)
    mod.test_full_size_decode_every_sample_of_every_fourth_stream(eng, native)
    mod.test_full_width_encode_every_slice(eng, native, "music", 4096)
    mod.test_full_width_encode_every_slice(eng, native, "noise", 512)
    for ch in (1, 3, 6, 7):
        mod.test_long_streams_every_sample_other_channel_counts(eng, native, ch)
    for ch, kind in ((1, "music"), (3, "noise"), (8, "music")):
        mod.test_encode_every_slice_other_channel_counts(eng, native, ch, kind)

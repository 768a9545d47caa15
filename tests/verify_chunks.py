"""Chunked drivers of oracle/vector_verify.py shared by the CPU tests (small batches, CPU tensors) and the GPU tests
(full-size batches, CUDA tensors): a batch is verified a few streams at a time so that the verifier's 64-bit
temporaries stay bounded."""
import torch

from oracle import vector_verify as vv

COUNT_KEYS = ("streams", "samples", "bad_file_headers", "bad_frame_headers", "bad_samples", "bad_streams", "boundaries",
              "chain_history_bad", "chain_weights_bad", "chain_bad_streams", "truncated_boundaries", "slices",
              "bad_slices", "unverifiable_slices", "bad_initial_state")


def _add(total, part):
    for k in COUNT_KEYS:
        if k in part:
            total[k] = total.get(k, 0) + part[k]
    return total


def decode_in_chunks(qoa, pcm, ch, n, rate=None, streams=None, chunk=8):
    """qoa [S, stride] uint8, pcm [S, >= n * ch] int16 (rows = streams; either device).  streams: indices to verify
    (default all).  Returns the summed counts of vv.verify_decode."""
    idx = list(range(qoa.shape[0])) if streams is None else list(streams)
    total = {}
    for b in range(0, len(idx), chunk):
        sel = torch.as_tensor(idx[b:b + chunk], dtype=torch.long, device=qoa.device)
        _add(total, vv.verify_decode(qoa.index_select(0, sel), pcm.index_select(0, sel)[:, :n * ch], ch, n, rate))
    return total


def encode_in_chunks(qoa, src, dec, ch, n, rate=None, chunk=64, max_lanes=1 << 22):
    """qoa [S, stride] uint8 = encoder output for src [S, n * ch]; dec [S, >= n * ch] = a decode of qoa.  Returns the summed
    counts of vv.verify_encode."""
    total = {}
    for b in range(0, qoa.shape[0], chunk):
        e = min(qoa.shape[0], b + chunk)
        _add(total, vv.verify_encode(qoa[b:e], src[b:e, :n * ch], dec[b:e, :n * ch], ch, n, rate, max_lanes=max_lanes))
    return total

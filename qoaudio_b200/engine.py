"""Engine: one context per GPU plus the batched host- and device-level calls over the C ABI.

This is the host-side mirror of the reference's encoder/decoder objects for BATCHES of streams.  The per-stream,
crate-shaped API (encode_all, QoaEncoder, QoaDecoder, ...) lives in api.py and is a thin layer over this module.
Per-stream tables (offsets, lengths, capacities) may be python sequences or numpy arrays; numpy arrays of the right
dtype are passed to the library without conversion (what bench.py uses for 8192-stream batches).
"""
import ctypes as C
import os
import threading

import numpy as np

from . import _native as N

_NP_OF = {C.c_uint64: np.uint64, C.c_size_t: np.uint64, C.c_uint32: np.uint32, C.c_int32: np.int32}

INFO_DTYPE = np.dtype([("status", np.int32), ("streaming", np.int32), ("header_samples", np.uint32),
                       ("channels", np.uint32), ("sample_rate", np.uint32), ("index_frames", np.uint32),
                       ("samples", np.uint64), ("frames", np.uint64), ("max_channels", np.uint32),
                       ("flags", np.uint32)])
assert INFO_DTYPE.itemsize == C.sizeof(N.StreamInfo)


class _Keep(list):
    """Builds ctypes array arguments and keeps their backing storage alive for the duration of a call."""

    def arr(self, values, ctype):
        if isinstance(values, np.ndarray):
            a = np.ascontiguousarray(values, dtype=_NP_OF[ctype])
            self.append(a)
            return C.cast(C.c_void_p(a.ctypes.data), C.POINTER(ctype))
        a = (ctype * len(values))(*values)
        self.append(a)
        return a


def _descs(descs):
    return (N.Desc * len(descs))(*[N.Desc(int(c), int(r), int(s)) for (c, r, s) in descs])


class Engine:
    """Owns a qoa_b200_ctx on one CUDA device."""

    def __init__(self, device=None):
        if device is None:
            device = int(os.environ.get("QOA_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        self.device = device
        self._h = C.c_void_p()
        rc = N.lib().qoa_b200_ctx_create(device, C.byref(self._h))
        if rc != 0:
            msg = N.lib().qoa_b200_last_error(None)
            raise N.NativeError(f"qoa_b200_ctx_create(device={device}) failed with {rc}: {msg.decode()} "
                                "(qoaudio_b200 has no CPU fallback)")

    def close(self):
        if self._h:
            N.lib().qoa_b200_ctx_destroy(self._h)
            self._h = C.c_void_p()

    # the native entry points behind encode_all_batch / decode_all_batch (MultiEngine swaps in the multi-device ones)
    def _encode_all_fn(self):
        return N.lib().qoa_b200_encode_all_batch

    def _decode_all_fn(self):
        return N.lib().qoa_b200_decode_all_batch

    def _check(self, rc, what):
        N.check(self._h, rc, "qoa_b200_" + what)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def kernel_launches(self):
        return int(N.lib().qoa_b200_kernel_launches(self._h))

    # ---- encode, host buffers ------------------------------------------------------------------------------
    def encode_all_batch(self, pcm_list, descs):
        """Batched encode_all (src/lib.rs:760).  pcm_list[i]: int16 array (interleaved), descs[i]: (channels,
        sample_rate, samples).  Returns (status list, list of np.uint8 arrays)."""
        n = len(pcm_list)
        assert len(descs) == n
        if n == 0:
            return [], []
        k = _Keep()
        pcm = [np.ascontiguousarray(p, dtype=np.int16).reshape(-1) for p in pcm_list]
        outs = []
        for (c, r, s) in descs:
            ok = 1 <= c <= 8 and r > 0 and s > 0
            outs.append(np.empty(N.lib().qoa_b200_encoded_size(c, s) if ok else 0, dtype=np.uint8))
        out_len = (C.c_size_t * n)()
        status = (C.c_int32 * n)()
        rc = self._encode_all_fn()(
            self._h, _descs(descs), k.arr([p.ctypes.data if p.size else None for p in pcm], C.c_void_p),
            k.arr([p.size for p in pcm], C.c_size_t), k.arr([o.ctypes.data if o.size else None for o in outs], C.c_void_p),
            k.arr([o.size for o in outs], C.c_size_t), out_len, status, n)
        self._check(rc, "encode_all_batch")
        return list(status), [o[:out_len[i]] for i, o in enumerate(outs)]

    # ---- encode, device buffers -----------------------------------------------------------------------------
    def encode_all_batch_dev(self, descs, d_pcm, pcm_off, pcm_len, d_out, out_off, out_cap, stream=0):
        """Device-resident batched encode_all.  d_pcm/d_out: raw device pointers (int).  Returns (status, out_len)."""
        n = len(descs)
        k = _Keep()
        out_len = (C.c_size_t * n)()
        status = (C.c_int32 * n)()
        rc = N.lib().qoa_b200_encode_all_batch_dev(
            self._h, _descs(descs), C.c_void_p(d_pcm), k.arr(pcm_off, C.c_uint64), k.arr(pcm_len, C.c_size_t),
            C.c_void_p(d_out), k.arr(out_off, C.c_uint64), k.arr(out_cap, C.c_size_t), out_len, status, n,
            C.c_void_p(stream))
        N.check(self._h, rc, "qoa_b200_encode_all_batch_dev")
        return list(status), list(out_len)

    def encoder_batch(self, descs):
        return EncoderBatch(self, descs)

    # ---- decode -----------------------------------------------------------------------------------------------
    def decode_plan(self, streams, mode=N.DECODE_ITER):
        """Index a batch of host-resident .qoa byte strings."""
        return DecodePlan._from_host(self, streams, mode)

    def decode_plan_dev(self, d_qoa, in_off, lens, mode=N.DECODE_ITER, stream=0):
        return DecodePlan._from_dev(self, d_qoa, in_off, lens, mode, stream)

    def decode_batch(self, streams, mode=N.DECODE_ITER):
        """Decode host-resident streams.  Returns (infos, list of np.int16 arrays)."""
        plan = self.decode_plan(streams, mode)
        try:
            return plan.info, plan.run()
        finally:
            plan.close()

    def decode_all_batch(self, streams, caps, mode=N.DECODE_ALL, out=None):
        """One-call host path (qoa_b200_decode_all_batch): large batches are pipelined over chunks of streams so that
        H2D, decode and D2H overlap.  streams: list of np.uint8 arrays; caps: capacity (i16 values) per stream.
        out (optional): C-contiguous int16 array of shape (len(streams), >= max(caps)) to decode into, one row per
        stream -- rows of one array (on either side) move as a single strided copy instead of one copy per stream.
        Returns (info structured array, status np.int32 array, list of np.int16 arrays of length caps[i])."""
        n = len(streams)
        bufs = [np.ascontiguousarray(s, dtype=np.uint8) for s in streams]
        if out is not None:
            assert out.dtype == np.int16 and out.ndim == 2 and out.flags.c_contiguous and out.shape[0] == n
            assert n == 0 or out.shape[1] >= max(int(c) for c in caps)
            outs = [out[i, :int(c)] for i, c in enumerate(caps)]
        else:
            outs = [np.empty(int(c), dtype=np.int16) for c in caps]
        info = np.zeros(n, dtype=INFO_DTYPE)
        status = np.zeros(n, dtype=np.int32)
        k = _Keep()
        rc = self._decode_all_fn()(
            self._h, k.arr([b.ctypes.data if b.size else None for b in bufs], C.c_void_p),
            k.arr([b.size for b in bufs], C.c_size_t), n, mode,
            k.arr([o.ctypes.data if o.size else None for o in outs], C.c_void_p), k.arr([o.size for o in outs], C.c_size_t),
            C.c_void_p(info.ctypes.data), C.c_void_p(status.ctypes.data))
        self._check(rc, "decode_all_batch")
        return info, status, outs

    # ---- generators / microbenchmarks --------------------------------------------------------------------------
    def gen_pcm_dev(self, d_pcm, n_streams, first_stream, channels, samples, start, seed, kind="music", stream=0):
        rc = N.lib().qoa_b200_gen_pcm_dev(self._h, C.c_void_p(d_pcm), n_streams, first_stream, channels, samples, start,
                                          seed, 1 if kind == "noise" else 0, C.c_void_p(stream))
        N.check(self._h, rc, "qoa_b200_gen_pcm_dev")

    def gen_qoa_dev(self, d_qoa, stride, n_streams, first_stream, channels, sample_rate, samples, seed, stream=0):
        rc = N.lib().qoa_b200_gen_qoa_dev(self._h, C.c_void_p(d_qoa), stride, n_streams, first_stream, channels,
                                          sample_rate, samples, seed, C.c_void_p(stream))
        N.check(self._h, rc, "qoa_b200_gen_qoa_dev")

    def int_issue_peak(self, kind):
        v = C.c_double(0)
        rc = N.lib().qoa_b200_int_issue_peak(self._h, kind, C.byref(v))
        N.check(self._h, rc, "qoa_b200_int_issue_peak")
        return v.value


class EncoderBatch:
    """Batch of streaming encoders: QoaEncoder::{new, write_header, encode_frame} (src/lib.rs:335, :404, :418)."""

    def __init__(self, engine, descs):
        self.engine = engine
        self.descs = [(int(c), int(r), int(s)) for (c, r, s) in descs]
        n = len(self.descs)
        st = (C.c_int32 * n)()
        self._h = C.c_void_p()
        rc = N.lib().qoa_b200_encoder_batch_new(engine.handle, _descs(self.descs), n, st, C.byref(self._h))
        N.check(engine.handle, rc, "qoa_b200_encoder_batch_new")
        self.status = list(st)

    def close(self):
        if self._h:
            N.lib().qoa_b200_encoder_batch_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def write_header(self):
        n = len(self.descs)
        bufs = [np.zeros(8, dtype=np.uint8) for _ in range(n)]
        k = _Keep()
        rc = N.lib().qoa_b200_encoder_batch_write_header(self._h, k.arr([b.ctypes.data for b in bufs], C.c_void_p))
        N.check(self.engine.handle, rc, "qoa_b200_encoder_batch_write_header")
        return [b.tobytes() for b in bufs]

    def encode_frame(self, pcm_list):
        """pcm_list[i]: int16 array for stream i's next frame, or None to skip the stream this round.
        Returns (status list, list of bytes)."""
        return self._encode_host(pcm_list, N.lib().qoa_b200_encoder_batch_encode_frame, False)

    def encode_chunk(self, pcm_list):
        """Like encode_frame but any number of samples per stream (cut into frames of 5120)."""
        return self._encode_host(pcm_list, N.lib().qoa_b200_encoder_batch_encode_chunk, True)

    def _encode_host(self, pcm_list, fn, chunk):
        n = len(self.descs)
        assert len(pcm_list) == n
        pcm = [None if p is None else np.ascontiguousarray(p, dtype=np.int16).reshape(-1) for p in pcm_list]
        if chunk:
            caps = [8 if p is None or c == 0 else N.lib().qoa_b200_encoded_size(c, -(-p.size // c))
                    for p, (c, _, _) in zip(pcm, self.descs)]
        else:
            caps = [8 + 16 * 8 + 8 * 256 * 8] * n
        outs = [np.empty(c, dtype=np.uint8) for c in caps]
        written = (C.c_size_t * n)()
        status = (C.c_int32 * n)()
        dummy = np.zeros(1, dtype=np.int16)  # a non-NULL pointer for "fed an empty frame" (-> InvalidSamples)
        ptrs = [None if p is None else (p.ctypes.data if p.size else dummy.ctypes.data) for p in pcm]
        k = _Keep()
        rc = fn(self._h, k.arr(ptrs, C.c_void_p), k.arr([0 if p is None else p.size for p in pcm], C.c_size_t),
                k.arr([o.ctypes.data for o in outs], C.c_void_p), k.arr(caps, C.c_size_t), written, status)
        N.check(self.engine.handle, rc, "qoa_b200_encoder_batch_encode_frame/chunk")
        return list(status), [outs[i][:written[i]].tobytes() for i in range(n)]

    def encode_chunk_dev(self, d_pcm, pcm_off, chunk_samples, d_out, out_off, stream=0):
        k = _Keep()
        rc = N.lib().qoa_b200_encoder_batch_encode_chunk_dev(
            self._h, C.c_void_p(d_pcm), k.arr(pcm_off, C.c_uint64), k.arr(chunk_samples, C.c_uint32),
            C.c_void_p(d_out), k.arr(out_off, C.c_uint64), C.c_void_p(stream))
        N.check(self.engine.handle, rc, "qoa_b200_encoder_batch_encode_chunk_dev")

    def lms(self, i, c):
        h = (C.c_int32 * 4)()
        w = (C.c_int32 * 4)()
        rc = N.lib().qoa_b200_encoder_batch_get_lms(self._h, i, c, h, w)
        N.check(self.engine.handle, rc, "qoa_b200_encoder_batch_get_lms")
        return list(h), list(w)


class DecodePlan:
    """Frame index of a batch (the validation pre-pass) + decode."""

    def __init__(self):
        self._h = C.c_void_p()
        self.engine = None
        self.n = 0
        self.info_array = np.zeros(0, dtype=INFO_DTYPE)
        self._info = None
        self._keep = None

    def _load_info(self):
        self.info_array = np.zeros(self.n, dtype=INFO_DTYPE)
        if self.n:
            N.lib().qoa_b200_decode_plan_info(self._h, C.c_void_p(self.info_array.ctypes.data))

    @property
    def info(self):
        """Per-stream dicts (status, streaming, header_samples, channels, sample_rate, samples, frames, ...)."""
        if self._info is None:
            names = INFO_DTYPE.names
            self._info = [{k: (bool(row[k]) if k == "streaming" else int(row[k])) for k in names}
                          for row in self.info_array]
        return self._info

    @classmethod
    def _from_host(cls, engine, streams, mode):
        self = cls()
        self.engine = engine
        bufs = [np.frombuffer(bytes(s), dtype=np.uint8) if not isinstance(s, np.ndarray)
                else np.ascontiguousarray(s, dtype=np.uint8) for s in streams]
        n = len(bufs)
        self.n = n
        self._keep = bufs
        k = _Keep()
        rc = N.lib().qoa_b200_decode_plan_create(
            engine.handle, k.arr([b.ctypes.data if b.size else None for b in bufs], C.c_void_p),
            k.arr([b.size for b in bufs], C.c_size_t), n, mode, C.byref(self._h))
        N.check(engine.handle, rc, "qoa_b200_decode_plan_create")
        self._load_info()
        return self

    @classmethod
    def _from_dev(cls, engine, d_qoa, in_off, lens, mode, stream):
        self = cls()
        self.engine = engine
        n = len(lens)
        self.n = n
        k = _Keep()
        rc = N.lib().qoa_b200_decode_plan_create_dev(engine.handle, C.c_void_p(d_qoa), k.arr(in_off, C.c_uint64),
                                                     k.arr(lens, C.c_size_t), n, mode, C.c_void_p(stream),
                                                     C.byref(self._h))
        N.check(engine.handle, rc, "qoa_b200_decode_plan_create_dev")
        self._load_info()
        return self

    def close(self):
        if self._h:
            N.lib().qoa_b200_decode_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def frame_headers(self, i):
        cnt = int(self.info_array["index_frames"][i])
        arr = (N.FrameHeaderC * max(cnt, 1))()
        got = N.lib().qoa_b200_decode_plan_frame_headers(self._h, i, arr, cnt)
        return [(a.num_channels, a.sample_rate, a.num_samples_per_channel) for a in arr[:min(got, cnt)]]

    def run(self, caps=None):
        """Decode into fresh host arrays.  Returns list of np.int16 arrays (what the iterator yields as samples)."""
        n = self.n
        if n == 0:
            return []
        counts = [int(v) for v in self.info_array["samples"]]
        if caps is None:
            caps = counts
        outs = [np.empty(c, dtype=np.int16) for c in caps]
        status = (C.c_int32 * n)()
        k = _Keep()
        rc = N.lib().qoa_b200_decode_plan_run(self._h, k.arr([o.ctypes.data if o.size else None for o in outs], C.c_void_p),
                                              k.arr(caps, C.c_size_t), status)
        N.check(self.engine.handle, rc, "qoa_b200_decode_plan_run")
        self.run_status = list(status)
        return [o[:min(counts[i], caps[i])] if status[i] != N.DEC_OUTPUT_TOO_SMALL else o[:0]
                for i, o in enumerate(outs)]

    def run_dev(self, d_out, out_off, out_cap, stream=0):
        """Decode into a device arena; returns the per-stream status as an np.int32 array."""
        n = self.n
        status = np.zeros(n, dtype=np.int32)
        k = _Keep()
        rc = N.lib().qoa_b200_decode_plan_run_dev(self._h, C.c_void_p(d_out), k.arr(out_off, C.c_uint64),
                                                  k.arr(out_cap, C.c_size_t), C.c_void_p(status.ctypes.data),
                                                  C.c_void_p(stream))
        N.check(self.engine.handle, rc, "qoa_b200_decode_plan_run_dev")
        return status


_default = None
_lock = threading.Lock()


class MultiEngine(Engine):
    """Owns a qoa_b200_mctx: several GPUs driven from this process.  encode_all_batch / decode_all_batch have the
    single-device signatures and results; the batch is cut into contiguous ranges of streams, one per device, balanced
    by work (qoa_b200_partition), no collective.  `engine(i)` exposes slot i's single-device context for the *_dev
    calls."""

    def __init__(self, devices=None):
        self._h = C.c_void_p()
        arr = (C.c_int * len(devices))(*devices) if devices else None
        rc = N.lib().qoa_b200_ctx_create_multi(arr, len(devices) if devices else 0, C.byref(self._h))
        if rc != 0:
            msg = N.lib().qoa_b200_mctx_last_error(None)
            raise N.NativeError(f"qoa_b200_ctx_create_multi({devices}) failed with {rc}: {msg.decode()} "
                                "(qoaudio_b200 has no CPU fallback)")
        self.devices = [N.lib().qoa_b200_mctx_device(self._h, i) for i in range(N.lib().qoa_b200_mctx_device_count(self._h))]
        self.device = self.devices[0]

    def close(self):
        if self._h:
            N.lib().qoa_b200_mctx_destroy(self._h)
            self._h = C.c_void_p()

    def kernel_launches(self):
        return int(N.lib().qoa_b200_mctx_kernel_launches(self._h))

    def _encode_all_fn(self):
        return N.lib().qoa_b200_mctx_encode_all_batch

    def _decode_all_fn(self):
        return N.lib().qoa_b200_mctx_decode_all_batch

    def _check(self, rc, what):
        if rc != 0:
            msg = N.lib().qoa_b200_mctx_last_error(self._h)
            raise N.NativeError(f"qoa_b200_mctx_{what} failed with {rc}: {msg.decode() if msg else '?'}")

    def decode_batch(self, streams, mode=N.DECODE_ITER):
        """Decode host-resident streams of unknown size on every device: sizes come from qoa_b200_probe (host-side header
        walk), the samples from the sharded one-call path.  Returns (infos, list of np.int16 arrays) like Engine."""
        bufs = [np.frombuffer(bytes(s), dtype=np.uint8) if not isinstance(s, np.ndarray) else s for s in streams]
        infos = [probe(b, mode) for b in bufs]
        info, status, outs = self.decode_all_batch(bufs, [i["samples"] for i in infos], mode)
        names = INFO_DTYPE.names
        return [{k: (bool(r[k]) if k == "streaming" else int(r[k])) for k in names} for r in info], outs

    def partition(self, weights):
        """The contiguous split the batch calls use: [(begin, end), ...] per device."""
        return partition(weights, len(self.devices))

    def engine(self, i):
        """A borrowed single-device Engine view of slot i (do not close it)."""
        e = Engine.__new__(Engine)
        e.device = self.devices[i]
        e._h = C.c_void_p(N.lib().qoa_b200_mctx_ctx(self._h, i))
        e.close = lambda: None
        return e


def probe(stream, mode=N.DECODE_ITER):
    """qoa_b200_probe: per-stream info dict (status, channels, sample_rate, samples, frames, ...) from a host-side walk
    of the frame headers (no GPU, no sample decoded) - what a caller needs to size the output of a one-call decode."""
    buf = np.frombuffer(bytes(stream), dtype=np.uint8) if not isinstance(stream, np.ndarray) else np.ascontiguousarray(stream, dtype=np.uint8)
    info = np.zeros(1, dtype=INFO_DTYPE)
    rc = N.lib().qoa_b200_probe(C.c_void_p(buf.ctypes.data if buf.size else None), buf.size, mode, C.c_void_p(info.ctypes.data))
    if rc != 0:
        raise ValueError(f"qoa_b200_probe failed with {rc}")
    return {k: (bool(info[0][k]) if k == "streaming" else int(info[0][k])) for k in INFO_DTYPE.names}


def partition(weights, parts):
    """qoa_b200_partition: contiguous ranges of near-equal summed weight, [(begin, end), ...] (host arithmetic only)."""
    w = np.ascontiguousarray(weights, dtype=np.uint64)
    first = np.zeros(parts + 1, dtype=np.uint64)
    rc = N.lib().qoa_b200_partition(C.c_void_p(w.ctypes.data), w.size, parts, C.c_void_p(first.ctypes.data))
    if rc != 0:
        raise ValueError(f"qoa_b200_partition failed with {rc}")
    return [(int(first[p]), int(first[p + 1])) for p in range(parts)]


def default_engine():
    """Process-wide engine on device QOA_B200_DEVICE / LOCAL_RANK / 0 (one process per GPU)."""
    global _default
    with _lock:
        if _default is None:
            _default = Engine()
        return _default

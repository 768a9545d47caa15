#!/usr/bin/env python
"""bench.py — batch QOA decode/encode throughput on B200 (BASELINE.json metric: Msamples/s + roofline fraction).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One JSON line on stdout (rank 0).  Headline workload (config.workload) = BASELINE.json configs[2]:
batch decode of 8192 synthetic stereo 44.1 kHz 60 s .qoa streams PER GPU (weak scaling: streams are independent,
sharded by stream, no collective).  A "step" = index pre-pass + frame-parallel decode of the whole resident batch.
The same line carries an "encode" object: configs[3]'s shape (4096 stereo 48 kHz streams per GPU) encoded in time
chunks with the LMS state persistent on the device, a step = one chunk of --enc-frames frames of every stream.

  value        device-resident throughput (inputs in HBM, CUDA events on the launching stream, max over ranks)
  e2e          the same metric through the C-ABI call with HOST (pinned) buffers, H2D + D2H inside the timed region
  roofline     decode kernel: algorithmic bytes / kernel time vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline the CPU oracle (a C restatement of the reference crate; the crate itself needs rustc, absent here),
               one stream per thread on the host cores, on a bounded sample of the same workload
`--impl reference` times that CPU oracle alone, with all host threads, on the same config/metric.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

DEC_CH, DEC_RATE, DEC_SAMPLES = 2, 44100, 2646000          # configs[2]: 60 s stereo 44.1 kHz
ENC_CH, ENC_RATE, ENC_SAMPLES = 2, 48000, 14400000         # configs[3]: 5 min stereo 48 kHz
ENC_OPS_PER_SAMPLE = 944                                   # 16 candidates x 59 integer ops


def enc_size(ch, n):
    return 8 + ((n + 5119) // 5120) * (8 + 16 * ch) + ((n + 19) // 20) * 8 * ch


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU during the timed region (pynvml, every 20 ms)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = str(e)

    NAMES = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
             0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
             0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.NAMES.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.02)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_decode_sample(threads):
    """One stream of the headline shape per thread: decode cost is data-independent, so the sample is built like the
    GPU's generator (valid headers + random payload) instead of spending minutes of CPU encode on it."""
    import numpy as np
    size = enc_size(DEC_CH, DEC_SAMPLES)
    rng = np.random.default_rng(1)
    proto = np.frombuffer(rng.bytes(size), dtype=np.uint8).copy()
    proto[0:8] = np.frombuffer(b"qoaf" + DEC_SAMPLES.to_bytes(4, "big"), dtype=np.uint8)
    fsz = 8 + 16 * DEC_CH + 2048 * DEC_CH
    frames = (DEC_SAMPLES + 5119) // 5120
    for f in range(frames):
        flen = min(5120, DEC_SAMPLES - f * 5120)
        fs = 8 + 16 * DEC_CH + 8 * DEC_CH * ((flen + 19) // 20)
        hdr = (DEC_CH << 56) | (DEC_RATE << 32) | (flen << 16) | fs
        proto[8 + f * fsz:16 + f * fsz] = np.frombuffer(hdr.to_bytes(8, "big"), dtype=np.uint8)
    return np.tile(proto, threads), size


def cpu_decode_step(batch, size, threads, rounds):
    """Decodes rounds x threads streams, one stream per thread; returns seconds."""
    from oracle import binding as ob
    t0 = time.perf_counter()
    for _ in range(rounds):
        rc, _ = ob.mt_decode(batch, threads, size, DEC_SAMPLES * DEC_CH, threads)
        assert rc == 0
    return time.perf_counter() - t0


def cpu_reference(args, want_encode=True):
    """The CPU arm: oracle (kind 'port'), one stream per thread, bounded sample of the decode (and encode) workload."""
    import numpy as np

    import signals
    from oracle import binding as ob
    cores = os.cpu_count() or 1
    threads = cores
    n_dec = max(512, threads * 2)  # ~18 core-seconds of CPU work
    batch, size = cpu_decode_sample(threads)
    rounds = max(1, n_dec // threads)
    out_stride = DEC_SAMPLES * DEC_CH
    cpu_decode_step(batch, size, threads, 1)  # warm-up (page faults)
    dt = cpu_decode_step(batch, size, threads, rounds)
    dec = {"value": rounds * threads * out_stride / dt / 1e6, "unit": "Msamples/s", "cores": threads, "kind": "port",
           "sample": f"{rounds * threads} streams of the headline shape (60 s stereo 44.1 kHz, synthetic payload), "
                     f"one stream per thread, {dt:.1f} s wall"}
    enc = None
    if want_encode:
        n = 5120 * 40  # 4.3 s of the 48 kHz stereo shape per stream
        pcm = np.concatenate([signals.synth_pcm(1, s, ENC_CH, n) for s in range(threads)])
        rc, _ = ob.mt_encode(pcm, threads, ENC_CH, ENC_RATE, n, threads)
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 8.0:
            rc, _ = ob.mt_encode(pcm, threads, ENC_CH, ENC_RATE, n, threads)
            assert rc == 0
            reps += 1
        dt = time.perf_counter() - t0
        enc = {"value": reps * threads * n * ENC_CH / dt / 1e6, "unit": "Msamples/s", "cores": threads, "kind": "port",
               "sample": f"{reps}x{threads} streams x {n} samples/ch stereo 48 kHz ('music' synthetic PCM, seed 1), "
                         f"reference search order with pruning, {dt:.1f} s wall"}
    return dec, enc


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (the oracle port: no Rust toolchain here), all
    host threads, W untimed + exactly K timed steps; a step is a bounded sample of the headline workload, sized so that
    the whole run stays within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, args.steps), max(0, args.warmup)
    threads = os.cpu_count() or 1
    batch, size = cpu_decode_sample(threads)
    per_stream = DEC_SAMPLES * DEC_CH
    t1 = cpu_decode_step(batch, size, threads, 1)  # page faults + a first estimate of one round
    rounds = max(1, min(512 // threads, int(120.0 / (steps + warm) / max(t1, 1e-3))))
    for _ in range(warm):
        cpu_decode_step(batch, size, threads, rounds)
    times = [cpu_decode_step(batch, size, threads, rounds) for _ in range(steps)]
    total = sum(times)
    v = steps * rounds * threads * per_stream / total / 1e6
    dec = {"value": v, "unit": "Msamples/s", "cores": threads, "kind": "port",
           "sample": f"{rounds * threads} streams of the headline shape (60 s stereo 44.1 kHz, synthetic payload) per step, "
                     f"one stream per thread, {steps} steps, {total:.1f} s wall"}
    line = {"impl": "reference", "metric": "batch decode throughput", "value": v, "unit": "Msamples/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": total / steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i32", "data": "synthetic",
            "config": workload_config(args), "cpu_baseline": dec,
            "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args):
    return {"workload": f"batch decode {args.streams} synthetic stereo 44.1 kHz 60 s .qoa streams per GPU "
                        "(BASELINE.json configs[2]), frame-parallel",
            "streams_per_gpu": args.streams, "channels": DEC_CH, "sample_rate": DEC_RATE,
            "samples_per_channel": DEC_SAMPLES, "l2": "inputs (2.1 MB x streams) larger than L2; no flush needed",
            "pcm_layout": "one slot per stream, slots start on multiples of one frame of PCM (20480 B)",
            "sharding": "streams across ranks, no collective"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=8192, help="decode streams per GPU")
    ap.add_argument("--enc-streams", type=int, default=4096, help="encode streams per GPU")
    ap.add_argument("--enc-frames", type=int, default=16, help="frames per stream per encode step (time chunk)")
    ap.add_argument("--e2e-streams", type=int, default=128)
    ap.add_argument("--no-encode", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-single", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    import qoaudio_b200 as q
    from qoaudio_b200.dist import Group
    from qoaudio_b200 import hostbind
    # pinned staging buffers of the e2e legs should live on the NUMA node next to this rank's GPU
    cpus_before, cpus_bound = hostbind.bind_to_gpu(local)
    dev = torch.device("cuda", local)
    grp = Group(device=dev)  # NCCL when WORLD_SIZE > 1: barrier + max/sum of timings only, no data-path collective
    eng = q.Engine(local)
    stream = torch.cuda.current_stream().cuda_stream
    W, K = max(args.warmup, 3), max(args.steps, 1)
    barrier, reduce_max = grp.barrier, grp.max

    # ---------------------------------------------------------------- decode (headline) -------------------------
    S = args.streams
    size = enc_size(DEC_CH, DEC_SAMPLES)
    stride = (size + 15) // 16 * 16
    out_per = DEC_SAMPLES * DEC_CH
    qoa = torch.empty(S * stride, dtype=torch.uint8, device=dev)
    pcm = torch.empty(S * ((out_per + 5120 * DEC_CH - 1) // (5120 * DEC_CH) * (5120 * DEC_CH)), dtype=torch.int16, device=dev)
    eng.gen_qoa_dev(qoa.data_ptr(), stride, S, rank * S, DEC_CH, DEC_RATE, DEC_SAMPLES, 1, stream)
    import numpy as np
    in_off = np.arange(S, dtype=np.uint64) * np.uint64(stride)
    lens = np.full(S, size, dtype=np.uint64)
    # every stream's PCM slot starts on a multiple of one full frame of PCM (20480 B): frames of consecutive streams
    # stay on the decode kernel's tensor-map row lattice; costs 0.04 % of extra address space
    out_slot = (out_per + 5120 * DEC_CH - 1) // (5120 * DEC_CH) * (5120 * DEC_CH)
    out_off = np.arange(S, dtype=np.uint64) * np.uint64(out_slot)
    caps = np.full(S, out_per, dtype=np.uint64)

    def decode_step(events=None):
        plan = eng.decode_plan_dev(qoa.data_ptr(), in_off, lens, q._native.DECODE_ALL, stream)
        if events is not None:
            events[0].record()
        st = plan.run_dev(pcm.data_ptr(), out_off, caps, stream)
        if events is not None:
            events[1].record()
        info = plan.info_array
        plan.close()
        return st, info

    for _ in range(W):
        st, info = decode_step()
    torch.cuda.synchronize()
    assert (st == 0).all() and (info["samples"] == out_per).all() and (info["status"] == 0).all()
    sampler = ClockSampler(local)
    kern_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = eng.kernel_launches()
    barrier()
    sampler.start()
    e0.record()
    for k in range(K):
        decode_step(kern_ev[k])
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = eng.kernel_launches() - launches0
    dec_ms = reduce_max(e0.elapsed_time(e1))
    kern_ms = statistics.mean(a.elapsed_time(b) for a, b in kern_ev)
    dec_samples_per_step = S * out_per * world
    dec_value = dec_samples_per_step * K / (dec_ms * 1e-3) / 1e6
    alg_bytes = S * (size + 2 * out_per)
    # spot parity inside the bench: first and last stream of this rank vs the oracle (checker only, untimed)
    parity = None
    if rank == 0 and not args.no_cpu:
        from oracle import binding as ob
        for s in (0, S - 1):
            blob = qoa[s * stride:s * stride + size].cpu().numpy().tobytes()
            got = pcm[s * out_slot:s * out_slot + out_per].cpu().numpy()
            _, ref = ob.decode_all(blob)
            assert np.array_equal(got, ref), f"bench decode of stream {s} differs from the oracle"
        parity = "streams 0 and last bit-exact vs oracle"

    # ---------------------------------------------------------------- e2e: host buffers through the C ABI -------
    e2e = None
    if not args.no_e2e:
        E = min(args.e2e_streams, S)
        h_in = qoa[:E * stride].cpu().pin_memory()
        h_out = torch.empty(E * out_per, dtype=torch.int16).pin_memory()
        import ctypes as C
        N = q._native
        in_ptrs = (C.c_void_p * E)(*[h_in.data_ptr() + i * stride for i in range(E)])
        in_lens = (C.c_size_t * E)(*([size] * E))
        out_ptrs = (C.c_void_p * E)(*[h_out.data_ptr() + 2 * i * out_per for i in range(E)])
        out_caps = (C.c_size_t * E)(*([out_per] * E))
        stat = (C.c_int32 * E)()

        def e2e_step():
            rc = N.lib().qoa_b200_decode_all_batch(eng.handle, in_ptrs, in_lens, E, N.DECODE_ALL, out_ptrs, out_caps,
                                                   None, stat)
            assert rc == 0 and all(s == 0 for s in stat)

        for _ in range(3):
            e2e_step()
        reps = max(3, min(K, 10))
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            e2e_step()
        torch.cuda.synchronize()
        dt = reduce_max(time.perf_counter() - t0)
        e2e = {"value": E * out_per * world * reps / dt / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": E * size,
               "d2h_bytes_per_step": E * out_per * 2, "streams_per_gpu": E, "ms_per_step": dt / reps * 1e3,
               "api": "qoa_b200_decode_all_batch (pinned host buffers -> H2D -> index + decode -> D2H)"}
        if rank == 0:
            assert torch.equal(h_out[:out_per], pcm[:out_per].cpu()), "e2e output differs from the device path"
        del h_in, h_out
    del pcm, qoa
    torch.cuda.empty_cache()

    # ---------------------------------------------------------------- encode (configs[3] shape, time-chunked) ----
    encode = None
    if not args.no_encode:
        ES, F = args.enc_streams, args.enc_frames
        chunk = F * 5120
        epcm = torch.empty(ES * chunk * ENC_CH, dtype=torch.int16, device=dev)
        fbytes = 8 + 16 * ENC_CH + 2048 * ENC_CH
        eout = torch.empty(ES * F * fbytes, dtype=torch.uint8, device=dev)
        batch = eng.encoder_batch([(ENC_CH, ENC_RATE, ENC_SAMPLES)] * ES)
        p_off = np.arange(ES, dtype=np.uint64) * np.uint64(chunk * ENC_CH)
        o_off = np.arange(ES, dtype=np.uint64) * np.uint64(F * fbytes)
        cs = np.full(ES, chunk, dtype=np.uint32)
        pos = 0
        for _ in range(W):
            eng.gen_pcm_dev(epcm.data_ptr(), ES, rank * ES, ENC_CH, chunk, pos, 1, "music", stream)
            batch.encode_chunk_dev(epcm.data_ptr(), p_off, cs, eout.data_ptr(), o_off, stream)
            pos += chunk
        KE = max(3, min(K, 10))
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(KE)]
        sampler2 = ClockSampler(local)
        l0 = eng.kernel_launches()
        barrier()
        sampler2.start()
        for k in range(KE):
            eng.gen_pcm_dev(epcm.data_ptr(), ES, rank * ES, ENC_CH, chunk, pos, 1, "music", stream)  # untimed input prep
            ev[k][0].record()
            batch.encode_chunk_dev(epcm.data_ptr(), p_off, cs, eout.data_ptr(), o_off, stream)
            ev[k][1].record()
            pos += chunk
        barrier()
        clocks2 = sampler2.stop()
        enc_ms = reduce_max(sum(a.elapsed_time(b) for a, b in ev))
        enc_samples = ES * chunk * ENC_CH * world * KE
        enc_value = enc_samples / (enc_ms * 1e-3) / 1e6
        sm_count = torch.cuda.get_device_properties(local).multi_processor_count
        max_mhz = clocks2.get("sm_max_mhz") or 1965
        nominal_peak = sm_count * 128 * max_mhz * 1e6
        measured = {k: eng.int_issue_peak(i) for i, k in enumerate(("imad", "lop3", "mix"))}
        per_gpu_ops = enc_value * 1e6 / world * ENC_OPS_PER_SAMPLE
        encode = {"metric": "batch encode throughput", "value": enc_value, "unit": "Msamples/s",
                  "ms_per_step": enc_ms / KE, "steps": KE, "gpu_launches": eng.kernel_launches() - l0,
                  "config": {"workload": f"batch encode {ES} synthetic stereo 48 kHz streams per GPU (BASELINE.json "
                                         f"configs[3] shape), time-chunked: {F} frames ({chunk} samples/ch) of every "
                                         "stream per step, LMS state persistent on the device, PCM generated on device",
                             "streams_per_gpu": ES, "frames_per_step": F},
                  "roofline": {"bound": "int32-issue", "achieved": per_gpu_ops / 1e12,
                               "peak": nominal_peak / 1e12, "unit": "Tops/s (944 algorithmic int-ops per sample)",
                               "frac": per_gpu_ops / nominal_peak,
                               "peak_kind": f"nominal {sm_count} SMs x 128 lanes x {max_mhz} MHz",
                               "measured_issue_peaks_tops": {k: v / 1e12 for k, v in measured.items()},
                               "frac_of_measured_mix_peak": per_gpu_ops / measured["mix"], "traffic": None},
                  "clocks": clocks2}
        batch.close()
        # e2e: the same shape through qoa_b200_encode_all_batch with pinned HOST buffers (time-chunked pipeline:
        # H2D of chunk t+1 || kernel of chunk t || D2H of chunk t-1); 24 frames per stream per call
        if not args.no_e2e:
            import ctypes as C
            N = q._native
            EF = 24
            per = EF * 5120 * ENC_CH
            osz = enc_size(ENC_CH, EF * 5120)
            tmp = torch.empty(ES * per, dtype=torch.int16, device=dev)
            eng.gen_pcm_dev(tmp.data_ptr(), ES, rank * ES, ENC_CH, EF * 5120, 0, 1, "music", stream)
            torch.cuda.synchronize()
            h_pcm = tmp.cpu().pin_memory()
            del tmp
            h_enc = torch.empty(ES * osz, dtype=torch.uint8).pin_memory()
            descs_c = (N.Desc * ES)(*[N.Desc(ENC_CH, ENC_RATE, EF * 5120)] * ES)
            pp = (C.c_void_p * ES)(*[h_pcm.data_ptr() + 2 * i * per for i in range(ES)])
            pl = (C.c_size_t * ES)(*([per] * ES))
            op_ = (C.c_void_p * ES)(*[h_enc.data_ptr() + i * osz for i in range(ES)])
            oc = (C.c_size_t * ES)(*([osz] * ES))
            ol = (C.c_size_t * ES)()
            stt = (C.c_int32 * ES)()

            def enc_e2e_step():
                rc = N.lib().qoa_b200_encode_all_batch(eng.handle, descs_c, pp, pl, op_, oc, ol, stt, ES)
                assert rc == 0 and all(v == 0 for v in stt)

            for _ in range(2):
                enc_e2e_step()
            barrier()
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                enc_e2e_step()
            torch.cuda.synchronize()
            dt = reduce_max(time.perf_counter() - t0)
            encode["e2e"] = {"value": ES * per * world * reps / dt / 1e6, "unit": "Msamples/s",
                             "h2d_bytes_per_step": ES * per * 2, "d2h_bytes_per_step": ES * osz, "ms_per_step": dt / reps * 1e3,
                             "api": "qoa_b200_encode_all_batch (pinned host buffers, 24 frames per stream per call, pipelined in time chunks of ~176 MB of PCM)"}
            del h_pcm, h_enc
        del epcm, eout

    # ---------------------------------------------------------------- configs[1]: ONE 54 s stereo stream --------
    # Latency-bound by construction (encode = one warp, decode = 466 threads): reported, not tuned for.
    single = None
    if rank == 0 and not args.no_single:
        SN, SCH, SRATE = 2381400, 2, 44100
        spcm = torch.empty(SN * SCH, dtype=torch.int16, device=dev)
        eng.gen_pcm_dev(spcm.data_ptr(), 1, 0, SCH, SN, 0, 1, "music", stream)
        ssize = enc_size(SCH, SN)
        sq = torch.zeros((ssize + 15) // 16 * 16, dtype=torch.uint8, device=dev)
        sout = torch.empty(SN * SCH, dtype=torch.int16, device=dev)
        t_enc, t_dec = [], []
        for it in range(3):
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record()
            st, ln = eng.encode_all_batch_dev([(SCH, SRATE, SN)], spcm.data_ptr(), [0], [SN * SCH], sq.data_ptr(), [0],
                                              [sq.numel()], stream)
            b.record()
            plan = eng.decode_plan_dev(sq.data_ptr(), [0], [ssize], q._native.DECODE_ALL, stream)
            plan.run_dev(sout.data_ptr(), [0], [SN * SCH], stream)
            c.record()
            torch.cuda.synchronize()
            plan.close()
            assert st == [0] and ln == [ssize]
            if it:
                t_enc.append(a.elapsed_time(b))
                t_dec.append(b.elapsed_time(c))
        single = {"workload": "single synthetic 54 s stereo 44.1 kHz stream (BASELINE.json configs[1])",
                  "encode_ms": min(t_enc), "decode_ms": min(t_dec),
                  "encode_msamples_s": SN * SCH / min(t_enc) / 1e3, "decode_msamples_s": SN * SCH / min(t_dec) / 1e3}
        if not args.no_cpu:
            from oracle import binding as ob
            hp = spcm.cpu().numpy()
            t0 = time.perf_counter()
            _, ref = ob.encode_all(hp, SCH, SRATE, SN)
            t1 = time.perf_counter()
            _, rpcm = ob.decode_all(ref)
            t2 = time.perf_counter()
            assert sq[:ssize].cpu().numpy().tobytes() == ref, "single-stream encode differs from the oracle"
            assert (sout.cpu().numpy() == rpcm).all(), "single-stream decode differs from the oracle"
            single.update({"cpu_encode_ms": (t1 - t0) * 1e3, "cpu_decode_ms": (t2 - t1) * 1e3,
                           "parity": "bit-exact vs oracle"})
        del spcm, sq, sout

    if rank != 0:
        grp.close()
        return

    cpu = cpu_enc = None
    if not args.no_cpu:
        hostbind.restore(cpus_before)  # the CPU baseline gets every host core
        cpu, cpu_enc = cpu_reference(args, want_encode=not args.no_encode)
        if encode is not None:
            encode["cpu_baseline"] = cpu_enc
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    traffic = None  # dram__bytes_read + dram__bytes_write of the decode kernel (ncu --set full), scaled per launch
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["decode_fast_kernel"]
        traffic = t["dram_bytes"] / t["streams"] * S
    except Exception:  # noqa: BLE001
        pass
    line = {
        "metric": "batch decode throughput", "value": dec_value, "unit": "Msamples/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": dec_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "i32", "data": "synthetic", "config": workload_config(args),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic,
                     "kernel": ("decode_fast_kernel<CH=2,S=3,NB=1,512,two codes per lookup>" if S * 517 >= 512 * 148 * 8
                                else "decode_fast_kernel<CH=2,S=3,NB=1,256>"), "kernel_ms": kern_ms,
                     "algorithmic_bytes_per_launch": alg_bytes,
                     "peak_kind": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "of fallback",
                     "note": "decode writes 8.5x what it reads; write-only HBM measured at 3.94 TB/s on this part "
                             "(tools/hbm_probe.py), i.e. a workload ceiling of ~0.72 of the copy peak"},
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "parity": parity,
        "encode": encode, "single_stream": single,
        "host_affinity": {"bound_cpus": len(cpus_bound) if cpus_bound else None, "allowed_cpus": len(cpus_before),
                          "note": "rank bound to its GPU's NUMA-local CPUs before pinned buffers are allocated; null = "
                                  "NVML reports no locality narrower than the allowed set (single-node host or VM)"},
    }
    print(json.dumps(line))
    grp.close()


if __name__ == "__main__":
    main()

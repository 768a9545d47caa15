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
    ap.add_argument("--streams", type=int, default=8192, help="decode streams per GPU (weak scaling) = total streams of the strong-scaling run")
    ap.add_argument("--enc-streams", type=int, default=4096, help="encode streams per GPU (weak) = total streams of the strong-scaling run")
    ap.add_argument("--enc-frames", type=int, default=16, help="frames per stream per encode step (time chunk)")
    ap.add_argument("--e2e-streams", type=int, default=256)
    ap.add_argument("--no-encode", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-single", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip noise / seeds / full-length / mixed-batch / strong-scaling legs")
    args = ap.parse_args()
    # A bench that hangs (a lost device, a rank that died before a barrier) must not hold the box until the driver's own
    # limit: after QOA_B200_BENCH_WATCHDOG_S seconds (default 40 min, far above any run) dump the stacks and exit.
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("QOA_B200_BENCH_WATCHDOG_S", "2400")), exit=True)
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    import qoaudio_b200 as q
    from qoaudio_b200.dist import Group
    from qoaudio_b200 import hostbind, shard
    # pinned staging buffers of the e2e legs should live on the NUMA node next to this rank's GPU
    cpus_before, cpus_bound = hostbind.bind_to_gpu(local)
    dev = torch.device("cuda", local)
    grp = Group(device=dev)  # NCCL when WORLD_SIZE > 1: barrier + max/sum of timings only, no data-path collective
    eng = q.Engine(local)
    stream = torch.cuda.current_stream().cuda_stream
    W, K = max(args.warmup, 3), max(args.steps, 1)
    barrier, reduce_max = grp.barrier, grp.max
    extras = not args.no_extras
    Ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def guarded(name, fn, collective=False):
        """The optional legs must not cost the headline line: a failure inside one is reported in that leg's object
        ({"error": ...}, loud in the JSON) instead of ending the run.  Legs with barriers in them are only guarded in a
        single-process run (a rank that skipped a barrier would leave the others waiting)."""
        if collective and world > 1:
            return fn()
        try:
            return fn()
        except Exception as e:  # noqa: BLE001
            import traceback
            sys.stderr.write(f"bench leg {name} failed:\n{traceback.format_exc()}\n")
            try:
                torch.cuda.synchronize()
            except Exception:  # noqa: BLE001
                pass
            return {"error": f"{type(e).__name__}: {e}"[:400]}

    # ---------------------------------------------------------------- decode (headline) -------------------------
    S = args.streams
    size = enc_size(DEC_CH, DEC_SAMPLES)
    stride = (size + 15) // 16 * 16
    out_per = DEC_SAMPLES * DEC_CH
    # every stream's PCM slot starts on a multiple of one full frame of PCM (20480 B): frames of consecutive streams
    # stay on the decode kernel's tensor-map row lattice; costs 0.04 % of extra address space
    out_slot = (out_per + 5120 * DEC_CH - 1) // (5120 * DEC_CH) * (5120 * DEC_CH)
    qoa = torch.empty(S * stride, dtype=torch.uint8, device=dev)
    pcm = torch.empty(S * out_slot, dtype=torch.int16, device=dev)
    eng.gen_qoa_dev(qoa.data_ptr(), stride, S, rank * S, DEC_CH, DEC_RATE, DEC_SAMPLES, 1, stream)
    in_off = np.arange(S, dtype=np.uint64) * np.uint64(stride)
    lens = np.full(S, size, dtype=np.uint64)
    out_off = np.arange(S, dtype=np.uint64) * np.uint64(out_slot)
    caps = np.full(S, out_per, dtype=np.uint64)

    def decode_step(n, events=None):
        """index pre-pass + decode of the first n resident streams"""
        plan = eng.decode_plan_dev(qoa.data_ptr(), in_off[:n], lens[:n], q._native.DECODE_ALL, stream)
        if events is not None:
            events[0].record()
        st = plan.run_dev(pcm.data_ptr(), out_off[:n], caps[:n], stream)
        if events is not None:
            events[1].record()
        info = plan.info_array
        plan.close()
        return st, info

    def decode_timed(n, steps):
        """W warm-up + `steps` timed steps over n streams per rank: (ms per step max over ranks, kernel ms, clocks, launches)"""
        for _ in range(W):
            st, info = decode_step(n)
        torch.cuda.synchronize()
        assert (st == 0).all() and (info["samples"] == out_per).all() and (info["status"] == 0).all()
        sampler = ClockSampler(local)
        kern_ev = [(Ev(), Ev()) for _ in range(steps)]
        e0, e1 = Ev(), Ev()
        launches0 = eng.kernel_launches()
        barrier()
        sampler.start()
        e0.record()
        for k in range(steps):
            decode_step(n, kern_ev[k])
        e1.record()
        barrier()
        clocks = sampler.stop()
        ms = reduce_max(e0.elapsed_time(e1)) / steps
        return ms, statistics.mean(a.elapsed_time(b) for a, b in kern_ev), clocks, eng.kernel_launches() - launches0

    dec_ms, kern_ms, clocks, launches = decode_timed(S, K)
    dec_value = S * out_per * world / (dec_ms * 1e-3) / 1e6
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
        if extras:
            # ... and EVERY sample of 64 streams spread over the batch against the decode recurrences, on the device, with
            # the vectorised verifier (oracle/vector_verify.py: checker only, torch kernels, untimed)
            def verify_leg():
                import verify_chunks as vc
                pick = sorted(set(int(v) for v in np.linspace(0, S - 1, min(S, 64))))
                r = vc.decode_in_chunks(qoa.view(S, stride), pcm.view(S, out_slot), DEC_CH, DEC_SAMPLES, DEC_RATE,
                                        streams=pick, chunk=8)
                assert r["bad_samples"] == 0 and r["bad_frame_headers"] == 0 and r["bad_file_headers"] == 0, r
                return f"; all {r['samples']} samples of {r['streams']} streams satisfy the decode recurrences (verified on the device)"
            v = guarded("decode_verify", verify_leg)
            parity += v if isinstance(v, str) else f"; device-side verifier leg failed: {v['error']}"
    # strong scaling: the SAME 8192-stream batch split over the ranks (BASELINE.json configs[2] "at 1/2/4/8 B200")
    strong = {"decode": None, "encode": None, "note": "total batch fixed (configs[2]/[3] as written), streams split by "
              "qoa_b200_partition over the ranks; the headline `value` is the weak-scaling number (batch per GPU fixed)"}
    if world == 1:
        strong["decode"] = {"value": dec_value, "unit": "Msamples/s", "streams_total": S, "streams_per_gpu": S,
                            "ms_per_step": dec_ms, "roofline_frac_per_gpu": None, "same_as": "headline (1 GPU)"}
    elif extras:
        b, e = shard.partition_contiguous([size] * S, world)[rank]
        n_s = e - b
        s_ms, s_kern, _, _ = decode_timed(n_s, K)
        peak_gbs = peak_hbm()
        strong["decode"] = {"value": S * out_per / (s_ms * 1e-3) / 1e6, "unit": "Msamples/s", "streams_total": S,
                            "streams_per_gpu": n_s, "ms_per_step": s_ms, "kernel_ms": s_kern,
                            "roofline_frac_per_gpu": n_s * (size + 2 * out_per) / (s_kern * 1e-3) / 1e9 / peak_gbs}

    # ---------------------------------------------------------------- e2e: host buffers through the C ABI -------
    e2e = None

    def e2e_leg():
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
        # The ceiling of that leg on THIS host: the same bytes as bare pinned cudaMemcpyAsync copies (one H2D and one
        # D2H stream per rank, all ranks at once, nothing else): what the host side of N ranks can move at best.
        s_h2d, s_d2h = torch.cuda.Stream(), torch.cuda.Stream()
        d_in = torch.empty(E * stride, dtype=torch.uint8, device=dev)

        def bare_step():
            with torch.cuda.stream(s_h2d):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s_d2h):
                h_out.copy_(pcm[:E * out_per], non_blocking=True)

        for _ in range(2):
            bare_step()
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            bare_step()
        torch.cuda.synchronize()
        bdt = reduce_max(time.perf_counter() - t0)
        gb = (E * size + E * out_per * 2) * world * reps / bdt / 1e9
        e2e.update({"bare_copy_ms_per_step": bdt / reps * 1e3, "host_ceiling_gbs": gb,
                    "host_ceiling_msamples_s": E * out_per * world * reps / bdt / 1e6,
                    "frac_of_host_ceiling": (E * out_per * world * reps / dt) / (E * out_per * world * reps / bdt),
                    "host_ceiling_note": "same H2D + D2H bytes per step as bare pinned cudaMemcpyAsync on two streams per "
                                         "rank, all ranks concurrently: aggregate host<->device GB/s of this box"})
        del h_in, h_out, d_in
        return e2e

    if not args.no_e2e:
        e2e = guarded("e2e", e2e_leg, collective=True)

    # ---------------------------------------------------------------- configs[4]: a shuffled mixed batch ---------
    mixed = None
    if rank == 0 and extras:
        mixed = guarded("mixed_batch", lambda: mixed_batch_leg(eng, qoa, pcm, stream))
    del pcm, qoa
    torch.cuda.empty_cache()

    # ---------------------------------------------------------------- encode (configs[3] shape, time-chunked) ----
    encode = None
    if not args.no_encode:
        ES, F = args.enc_streams, args.enc_frames
        chunk = F * 5120
        fbytes = 8 + 16 * ENC_CH + 2048 * ENC_CH
        epcm = torch.empty(ES * chunk * ENC_CH, dtype=torch.int16, device=dev)
        eout = torch.empty(ES * F * fbytes, dtype=torch.uint8, device=dev)
        p_off = np.arange(ES, dtype=np.uint64) * np.uint64(chunk * ENC_CH)
        o_off = np.arange(ES, dtype=np.uint64) * np.uint64(F * fbytes)
        KE = max(3, min(K, 10))
        sm_count = torch.cuda.get_device_properties(local).multi_processor_count

        def encode_run(n, kind, seed, warm, steps, first_stream, total_samples=None, keep=()):
            """Time-chunked encode of n streams: `warm` untimed + `steps` timed chunks of F frames (or, with
            total_samples, the whole stream incl. its partial last chunk).  PCM is generated on the device (untimed);
            CUDA events bracket every encode call.  keep: streams whose output bytes are collected on the host."""
            batch = eng.encoder_batch([(ENC_CH, ENC_RATE, ENC_SAMPLES)] * n)
            pos, ev, kept = 0, [], {s: [] for s in keep}
            todo = warm + steps if total_samples is None else (total_samples + chunk - 1) // chunk
            sampler = ClockSampler(local)
            l0 = eng.kernel_launches()
            for c in range(todo):
                if c == warm:
                    barrier()
                    sampler.start()
                cs_n = chunk if total_samples is None else min(chunk, total_samples - pos)
                eng.gen_pcm_dev(epcm.data_ptr(), n, first_stream, ENC_CH, cs_n, pos, seed, kind, stream)  # untimed input prep
                a, b = Ev(), Ev()
                a.record()
                batch.encode_chunk_dev(epcm.data_ptr(), np.arange(n, dtype=np.uint64) * np.uint64(cs_n * ENC_CH),
                                       np.full(n, cs_n, dtype=np.uint32), eout.data_ptr(), o_off[:n], stream)
                b.record()
                if c >= warm:
                    ev.append((a, b))
                nbytes = enc_size(ENC_CH, cs_n) - 8
                for s in keep:
                    kept[s].append(eout[int(o_off[s]):int(o_off[s]) + nbytes].cpu().numpy().tobytes())
                pos += cs_n
            barrier()
            clk = sampler.stop()
            ms = reduce_max(sum(a.elapsed_time(b) for a, b in ev))
            batch.close()
            return {"ms": ms, "chunks": len(ev), "samples": (pos - warm * chunk) * ENC_CH * n, "clocks": clk,
                    "launches": eng.kernel_launches() - l0, "kept": kept, "end": pos}

        def oracle_frames(stream_id, kind, seed, n_samples):
            """The reference's streaming encoder (LMS carried across frames, src/lib.rs:481-482) on the same PCM."""
            from oracle import binding as ob
            enc = ob.StreamEncoder(ENC_CH, ENC_RATE, ENC_SAMPLES)
            pcm_h = signals_synth(seed, stream_id, ENC_CH, n_samples, kind)
            out = []
            for f0 in range(0, n_samples, 5120):
                st, b = enc.encode_frame(pcm_h[f0 * ENC_CH:min(f0 + 5120, n_samples) * ENC_CH])
                assert st == 0
                out.append(b)
            return out

        import signals
        signals_synth = signals.synth_pcm
        verify = rank == 0 and not args.no_cpu
        main_run = encode_run(ES, "music", 1, W, KE, rank * ES, keep=(0, ES - 1) if verify else ())
        enc_value = main_run["samples"] * world / (main_run["ms"] * 1e-3) / 1e6
        enc_parity = None
        if verify:  # every chunk of streams 0 and last (warm-up included: the LMS state is carried) vs the oracle
            for s in (0, ES - 1):
                ref = b"".join(oracle_frames(rank * ES + s, "music", 1, main_run["end"]))
                assert b"".join(main_run["kept"][s]) == ref, f"bench encode of stream {s} differs from the oracle"
            enc_parity = f"streams 0 and last, all {W + KE} chunks ({main_run['end']} samples/ch, LMS carried) bit-exact vs oracle"
        max_mhz = main_run["clocks"].get("sm_max_mhz") or 1965
        nominal_peak = sm_count * 128 * max_mhz * 1e6
        measured = {k: eng.int_issue_peak(i) for i, k in enumerate(("imad", "lop3", "mix"))}
        per_gpu_ops = enc_value * 1e6 / world * ENC_OPS_PER_SAMPLE
        issue_pct = None
        try:
            issue_pct = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["encode_kernel"].get("issue_active_pct")
        except Exception:  # noqa: BLE001
            pass
        encode = {"metric": "batch encode throughput", "value": enc_value, "unit": "Msamples/s",
                  "ms_per_step": main_run["ms"] / KE, "steps": KE, "gpu_launches": main_run["launches"],
                  "config": {"workload": f"batch encode {ES} synthetic stereo 48 kHz streams per GPU (BASELINE.json "
                                         f"configs[3] shape), time-chunked: {F} frames ({chunk} samples/ch) of every "
                                         "stream per step, LMS state persistent on the device, PCM generated on device",
                             "streams_per_gpu": ES, "frames_per_step": F, "signal": "music", "seed": 1},
                  "roofline": {"bound": "int32-issue", "achieved": per_gpu_ops / 1e12,
                               "peak": nominal_peak / 1e12, "unit": "Tops/s (944 algorithmic int-ops per sample)",
                               "frac": per_gpu_ops / nominal_peak,
                               "peak_kind": f"nominal {sm_count} SMs x 128 lanes x {max_mhz} MHz",
                               "issue_active_pct_ncu": issue_pct,
                               "frac_note": "algorithmic ops (SURVEY §8d: 59 source-level ops x 16 candidates) over the "
                                            "nominal issue peak; the kernel issues 42.5 instructions per candidate step, "
                                            "so the issue slots actually used are issue_active_pct_ncu (static, from "
                                            "profiles/, ncu --set full)",
                               "measured_issue_peaks_tops": {k: v / 1e12 for k, v in measured.items()},
                               "frac_of_measured_mix_peak": per_gpu_ops / measured["mix"], "traffic": None},
                  "clocks": main_run["clocks"], "parity": enc_parity}
        def encode_extras():
            # white noise (worst case for the CPU's pruning; identical work for the GPU's lock-step search) and seeds 2, 3
            variants = {}
            for kind, seed in (("noise", 1), ("music", 2), ("music", 3)):
                r = encode_run(ES, kind, seed, 1, 3, rank * ES, keep=(0,) if verify else ())
                v = {"value": r["samples"] * world / (r["ms"] * 1e-3) / 1e6, "unit": "Msamples/s", "steps": 3}
                if verify:
                    assert b"".join(r["kept"][0]) == b"".join(oracle_frames(rank * ES, kind, seed, r["end"]))
                    v["parity"] = "stream 0 bit-exact vs oracle"
                variants[f"{kind}_seed{seed}"] = v
            encode["noise"] = variants.pop("noise_seed1")
            encode["seeds"] = variants
            # one full-length configs[3] pass: 2813 frames per stream, the 2560-sample tail included
            r = encode_run(ES, "music", 1, 0, 0, rank * ES, total_samples=ENC_SAMPLES, keep=(0, ES - 1) if verify else ())
            full = {"value": r["samples"] * world / (r["ms"] * 1e-3) / 1e6, "unit": "Msamples/s", "kernel_ms_total": r["ms"],
                    "chunks": r["chunks"], "frames_per_stream": (ENC_SAMPLES + 5119) // 5120, "samples_per_channel": ENC_SAMPLES,
                    "streams_per_gpu": ES, "gsamples": r["samples"] * world / 1e9, "clocks": r["clocks"]}
            if verify:
                from oracle import binding as ob
                for s in (0, ES - 1):
                    st, ref = ob.encode_all(signals_synth(1, rank * ES + s, ENC_CH, ENC_SAMPLES, "music"), ENC_CH, ENC_RATE, ENC_SAMPLES)
                    assert st == 0 and b"".join(r["kept"][s]) == bytes(ref)[8:], f"full-length encode of stream {s} differs from the oracle"
                full["parity"] = "streams 0 and last, all 2813 frames, byte-identical to the oracle's one-shot encode_all"
            encode["full_c4"] = full
            # strong scaling: the SAME 4096-stream batch split over the ranks
            if world == 1:
                strong["encode"] = {"value": enc_value, "unit": "Msamples/s", "streams_total": ES, "streams_per_gpu": ES,
                                    "ms_per_step": main_run["ms"] / KE, "same_as": "encode.value (1 GPU)"}
            else:
                b, e = shard.partition_contiguous([ENC_SAMPLES * ENC_CH] * ES, world)[rank]
                r = encode_run(e - b, "music", 1, W, KE, b)
                strong["encode"] = {"value": ES * chunk * ENC_CH * KE / (r["ms"] * 1e-3) / 1e6, "unit": "Msamples/s",
                                    "streams_total": ES, "streams_per_gpu": e - b, "ms_per_step": r["ms"] / KE,
                                    "kernel": "latency-mode encode kernel" if (e - b) <= 1184 else "throughput encode kernel"}
            return True

        if extras:
            ok = guarded("encode_extras", encode_extras, collective=True)
            if ok is not True:
                encode["extras_error"] = ok["error"]

        # e2e: the same shape through qoa_b200_encode_all_batch with pinned HOST buffers (time-chunked pipeline:
        # H2D of chunk t+1 || kernel of chunk t || D2H of chunk t-1); 24 frames per stream per call
        def encode_e2e():
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
            return encode["e2e"]

        if not args.no_e2e:
            encode["e2e"] = guarded("encode_e2e", encode_e2e, collective=True)
        del epcm, eout

    # ---------------------------------------------------------------- configs[1]: ONE 54 s stereo stream --------
    # Latency-bound by construction (encode = one warp, decode = 466 threads): reported, not tuned for.
    single = None

    def single_leg():
        SN, SCH, SRATE = 2381400, 2, 44100
        spcm = torch.empty(SN * SCH, dtype=torch.int16, device=dev)
        eng.gen_pcm_dev(spcm.data_ptr(), 1, 0, SCH, SN, 0, 1, "music", stream)
        ssize = enc_size(SCH, SN)
        sq = torch.zeros((ssize + 15) // 16 * 16, dtype=torch.uint8, device=dev)
        sout = torch.empty(SN * SCH, dtype=torch.int16, device=dev)
        t_enc, t_dec = [], []
        for it in range(3):
            a, b, c = (Ev() for _ in range(3))
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
        return single

    if rank == 0 and not args.no_single:
        single = guarded("single_stream", single_leg)

    if rank != 0:
        grp.close()
        return

    cpu = cpu_enc = None
    if not args.no_cpu:
        hostbind.restore(cpus_before)  # the CPU baseline gets every host core
        cpu, cpu_enc = cpu_reference(args, want_encode=not args.no_encode)
        if encode is not None:
            encode["cpu_baseline"] = cpu_enc
    peak = peak_hbm()
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    traffic = traffic_src = None  # dram__bytes_read + dram__bytes_write of the decode kernel (ncu --set full), scaled per launch
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["decode_fast_kernel"]
        traffic = t["dram_bytes"] / t["streams"] * S
        traffic_src = "static: " + t["source"]
    except Exception:  # noqa: BLE001
        pass
    # The bound that actually binds: per stereo sample pair the loop needs 16 IMAD (FMA-heavy pipe) and 17
    # ALU-pipe instructions; each pipe accepts one warp instruction per two cycles, so a warp needs >= 33 cycles per pair
    # with both pipes perfectly overlapped.  tools/core_probe.cu measures the same ladder without memory traffic.
    sm_count = torch.cuda.get_device_properties(local).multi_processor_count
    pairs_per_launch = S * DEC_SAMPLES
    floor_ms = pairs_per_launch / 32 / (sm_count * 4) * 33 / ((clocks.get("sm_max_mhz") or 1965) * 1e6) * 1e3
    int_pipes = {"bound": "int32 pipes (FMA-heavy + ALU, one warp instruction per two cycles each)",
                 "per_sample_pair": {"imad": 16, "alu": 17, "lds": 2}, "floor_cycles_per_warp_and_pair": 33,
                 "floor_ms_at_max_clock": floor_ms, "kernel_ms": kern_ms, "frac": floor_ms / kern_ms,
                 "probe": "tools/core_probe.cu: pair_core alone 19.7 ms, + table lookups 23.6 ms, full kernel 26.2 ms "
                          "at 1965 MHz (profiles/r02_core_probe.txt)"}
    line = {
        "metric": "batch decode throughput", "value": dec_value, "unit": "Msamples/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": dec_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "i32", "data": "synthetic", "config": workload_config(args),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "kernel": "decode_fast_kernel<CH=2,S=3,NB=1,256> (uniform-warp loop + general loop)", "kernel_ms": kern_ms,
                     "algorithmic_bytes_per_launch": alg_bytes,
                     "peak_kind": "MEASURED_PEAKS.json hbm_gbs (of measured)" if PEAKS.get("hbm_gbs") else "of fallback",
                     "note": "the kernel is bound by integer issue (16 IMAD + 12 ALU per stereo sample pair on two "
                             "half-rate pipes), not by HBM: DRAM ~48 % busy, traffic = 1.02 x algorithmic "
                             "(profiles/r02_decode_experiments.md)",
                     "int_pipe_roofline": int_pipes},
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "parity": parity,
        "scaling_strong": strong, "mixed_batch": mixed,
        "encode": encode, "single_stream": single,
        "host_affinity": {"bound_cpus": len(cpus_bound) if cpus_bound else None, "allowed_cpus": len(cpus_before),
                          "note": "rank bound to its GPU's NUMA-local CPUs before pinned buffers are allocated; null = "
                                  "NVML reports no locality narrower than the allowed set (single-node host or VM)"},
    }
    print(json.dumps(line))
    grp.close()


PEAKS = {}
try:
    PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:  # noqa: BLE001
    pass


def peak_hbm():
    """Roofline denominator: the driver-measured copy bandwidth of this pool's B200s, else the profiling guide's fallback."""
    return float(PEAKS.get("hbm_gbs", 6650.0))


def mixed_batch_leg(eng, qoa, pcm, stream):
    """BASELINE.json configs[4]: a shuffled batch of mono ... 8-channel streams of 1-60 s at 8-96 kHz with odd lengths
    (partial last frames and slices) decoded in one plan: every stream runs on its own class's kernel.  Reuses the
    resident arenas."""
    import numpy as np
    import torch
    import qoaudio_b200 as q
    from oracle import binding as ob
    n_streams = 2048
    rng = np.random.default_rng(5)
    chs = rng.choice([1, 2, 3, 4, 5, 6, 7, 8], size=n_streams, p=[0.2, 0.3, 0.05, 0.1, 0.05, 0.15, 0.05, 0.1])
    rates = rng.choice([8000, 11025, 16000, 22050, 32000, 44100, 48000, 96000], size=n_streams)
    ns = rng.integers(1, 61, size=n_streams) * 44100 + rng.integers(0, 5120, size=n_streams)
    sizes = np.array([enc_size(int(c), int(n)) for n, c in zip(ns, chs)], dtype=np.uint64)
    strides = (sizes + np.uint64(15)) // np.uint64(16) * np.uint64(16)
    in_off = np.concatenate([[0], np.cumsum(strides)[:-1]]).astype(np.uint64)
    caps = (ns * chs).astype(np.uint64)
    slots = (caps + np.uint64(7)) // np.uint64(8) * np.uint64(8)
    out_off = np.concatenate([[0], np.cumsum(slots)[:-1]]).astype(np.uint64)
    if int(strides.sum()) > qoa.numel() or int(slots.sum()) > pcm.numel():
        return None  # a reduced --streams run: the resident arenas are too small for this leg
    for i in range(n_streams):
        eng.gen_qoa_dev(qoa.data_ptr() + int(in_off[i]), int(strides[i]), 1, i, int(chs[i]), int(rates[i]), int(ns[i]), 3, stream)
    torch.cuda.synchronize()
    best = 1e9
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan = eng.decode_plan_dev(qoa.data_ptr(), in_off, sizes, q._native.DECODE_ALL, stream)
        st = plan.run_dev(pcm.data_ptr(), out_off, caps, stream)
        e1.record()
        torch.cuda.synchronize()
        plan.close()
        assert (st == 0).all()
        if it:
            best = min(best, e0.elapsed_time(e1))
    for i in sorted(set([int(np.argmax(chs == c)) for c in range(1, 9)] + [0, n_streams - 1])):  # one stream per class + first, last
        blob = qoa[int(in_off[i]):int(in_off[i]) + int(sizes[i])].cpu().numpy().tobytes()
        got = pcm[int(out_off[i]):int(out_off[i]) + int(caps[i])].cpu().numpy()
        assert np.array_equal(got, ob.decode_all(blob)[1]), f"mixed-batch stream {i} differs from the oracle"
    samples = int(caps.sum())
    return {"workload": f"{n_streams} shuffled 1..8-channel streams of 1-60 s (+ odd remainders) at 8-96 kHz with partial last "
                        "frames and slices (BASELINE.json configs[4] shape), index + decode in one plan, one launch per class",
            "value": samples / best / 1e3, "unit": "Msamples/s", "ms": best, "gsamples": samples / 1e9,
            "streams_per_channel_count": {str(c): int((chs == c).sum()) for c in range(1, 9)},
            "parity": "one stream per channel count + first and last stream bit-exact vs oracle"}


if __name__ == "__main__":
    main()

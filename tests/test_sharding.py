"""Host-side multi-GPU logic on CPU: partitioning and the world-size-2 gloo reductions bench.py relies on."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_covers_and_balances():
    from qoaudio_b200.shard import partition_contiguous
    rng = np.random.default_rng(3)
    for world in (1, 2, 4, 8):
        for n in (0, 1, 7, 8, 100, 8192):
            costs = rng.integers(1, 1000, size=n).tolist()
            parts = partition_contiguous(costs, world)
            assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            if n >= 8 * world:
                loads = [sum(costs[b:e]) for b, e in parts]
                assert max(loads) <= sum(costs) / world + 1000
    # uniform batches split evenly (the bench shape)
    assert partition_contiguous([5] * 8192, 8) == [(i * 1024, (i + 1) * 1024) for i in range(8)]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    import numpy as np
    import torch.distributed as dist
    from qoaudio_b200.dist import Group
    from qoaudio_b200.shard import my_range, encode_cost
    import signals
    from oracle import binding as ob
    g = Group(backend="gloo")
    shapes = [(1 + i % 3, 8000, 50 + 37 * i) for i in range(11)]
    costs = [encode_cost(c, n) for c, _, n in shapes]
    b, e = my_range(costs)
    # every rank "processes" its shard (the oracle stands in for the GPU on this CPU-only box)
    local = {{i: ob.encode_all(signals.synth_pcm(1, i, shapes[i][0], shapes[i][2]), *shapes[i])[1].hex() for i in range(b, e)}}
    gathered = [None] * g.world
    dist.all_gather_object(gathered, local)
    units = sum(shapes[i][0] * shapes[i][2] for i in range(b, e))
    secs = 1.0 + g.rank  # rank 1 is the slow one
    tp = g.throughput(units, secs)
    g.barrier()
    if g.rank == 0:
        merged = {{}}
        for d in gathered:
            assert not (set(d) & set(merged))
            merged.update(d)
        print(json.dumps({{"n": len(merged), "tp": tp, "keys": sorted(merged), "sha": merged}}))
    g.close()
""")


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=240) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    import json
    import signals
    from oracle import binding as ob
    res = json.loads(outs[0][0].strip().splitlines()[-1])
    shapes = [(1 + i % 3, 8000, 50 + 37 * i) for i in range(11)]
    assert res["n"] == 11 and res["keys"] == [str(i) for i in range(11)] or res["keys"] == list(range(11))
    total_units = sum(c * n for c, _, n in shapes)
    assert abs(res["tp"] - total_units / 2.0) < 1e-6  # all units / the slowest rank's 2.0 s
    for i, s in enumerate(shapes):
        want = ob.encode_all(signals.synth_pcm(1, i, s[0], s[2]), *s)[1].hex()
        assert res["sha"][str(i)] == want


def test_partition_is_the_c_abi_one():
    """shard.partition_contiguous is qoa_b200_partition (host arithmetic of the C ABI, no GPU needed): the split a
    multi-device context uses inside one process is the split torchrun ranks use."""
    import ctypes as C
    from qoaudio_b200 import _native as N, partition
    from qoaudio_b200.shard import partition_contiguous
    w = np.array([7, 1, 1, 1, 9, 2, 2, 5], dtype=np.uint64)
    first = (C.c_size_t * 4)()
    assert N.lib().qoa_b200_partition(C.c_void_p(w.ctypes.data), w.size, 3, first) == 0
    assert list(first) == [0, 3, 5, 8] and partition(w, 3) == [(0, 3), (3, 5), (5, 8)] == partition_contiguous(w.tolist(), 3)
    assert N.lib().qoa_b200_partition(C.c_void_p(w.ctypes.data), w.size, 0, first) == N.E_ARG
    # without a device the multi-device context fails loudly (no CPU fallback), it does not degrade to something else
    import torch
    if not torch.cuda.is_available():
        h = C.c_void_p()
        assert N.lib().qoa_b200_ctx_create_multi(None, 0, C.byref(h)) != 0 and not h

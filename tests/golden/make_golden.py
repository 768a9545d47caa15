"""Regenerates tests/golden/golden.json.

Provenance: the reference (a Rust crate) cannot be built or imported in this image (no rustc/cargo), so these
vectors are produced by the repo's CPU oracle (oracle/qoa_oracle.c).  The `appendix_b` block holds values that
were produced INDEPENDENTLY during the survey by a separate scratch restatement; the
script refuses to write the file unless the oracle reproduces every one of them, and the small KATs are also
re-derived with the independent pure-Python restatement (oracle/qoa_pyref.py).

The fixture julien_baker_sprained_ankle.qoa is the reference's own decode fixture
(/root/reference/fixtures/, used by src/lib.rs:996 and benches/qoa_benchmarks.rs:3), copied as test data.

Run from the repo root:  python tests/golden/make_golden.py

Pinning the oracle to the crate itself (needs a Rust toolchain, which this image lacks — rust/golden-dump/ is the
30-line program that runs the crate):
    python tests/golden/make_golden.py --export-inputs DIR      # KATs, synthetic streams, edge signals, the fixture
    cargo run --release --manifest-path rust/golden-dump/Cargo.toml -- DIR OUT
    python tests/golden/make_golden.py --check-against OUT      # crate bytes == oracle bytes == golden.json
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import binding as ob  # noqa: E402
from oracle import qoa_pyref as py  # noqa: E402
import signals  # noqa: E402

APPENDIX_B = {
    "fixture_sha256": "b8d822ffee42abe052dfaab00136e86c3c1e9eb6e86cd700867b61a9f45a3372",
    "fixture_decoded_pcm_sha256": "2ed7d9a42d13659a6ff51b04728499d99ac99686326b87f215fdeabe05522619",
    "fixture_reencode_sha256": "e9f87726aef5d602e248dc839ac7de5c570ad869419984f00274cde76f28c19e",
    "fixture_reencode_decoded_sha256": "04b8d8751976767ab6ea13d92a151a15ecf500e47153c9141c78746fd805538a",
    "kats": {
        "saw1000_mono": {"len": 432, "sha256": "b05c7a346d8e34a1b0a26e0d648d525b7e56e3543559eadd86b9605e24924e9f",
                         "first_slices": ["df986092034d34d3", "8cc8208201208048", "1c2800221108a28a"]},
        "silence21_mono": {"len": 48, "sha256": "d93e4ecb4538e4ba463b59b194cbfb719ee72eb0f4ef90858eef910d44dcd959",
                           "first_slices": ["00d10d0200051451", "0400000000000000"]},
        "mix45_3ch": {"len": 136, "sha256": "ddbba666ff82956b70b73bbf3f6623ab42d53f018613d58aabca9a80a80c21ce",
                      "first_slices": ["ff99249048041201", "8fbefbefbefbefbe", "9b01048047c0fc07"]},
        "lcg5121_8ch": {"len": 16728, "sha256": "0a5a96933f10f228c17d6aed7a418bfb6433686aff837b5ea98d4f7d93ad7464",
                        "first_slices": ["c1f6df1ff6dd6a37", "d1f65f8dfef359bf", "ffb7e37dfef4ebbe"]},
    },
}


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def main():
    g = {"appendix_b": APPENDIX_B, "provenance": "oracle/qoa_oracle.c; appendix_b from SURVEY.md Appendix B"}
    data = open(os.path.join(ROOT, "tests/golden/julien_baker_sprained_ankle.qoa"), "rb").read()
    assert sha(data) == APPENDIX_B["fixture_sha256"]
    info, pcm, hdrs = ob.decode_iter(data)
    assert info["status"] == 0 and sha(pcm.tobytes()) == APPENDIX_B["fixture_decoded_pcm_sha256"]
    st, enc = ob.encode_all(pcm, 2, 44100, 2394122)
    assert st == 0 and sha(enc) == APPENDIX_B["fixture_reencode_sha256"]
    info2, pcm2 = ob.decode_all(enc)
    assert sha(pcm2.tobytes()) == APPENDIX_B["fixture_reencode_decoded_sha256"]
    g["fixture"] = {"channels": 2, "sample_rate": 44100, "samples": 2394122, "frames": 468,
                    "last_frame_samples": 3082, "bytes": len(data)}

    for name, (sig, ch, rate, n) in signals.kat_signals().items():
        st, enc = ob.encode_all(sig, ch, rate, n)
        exp = APPENDIX_B["kats"][name]
        assert st == 0 and len(enc) == exp["len"] and sha(enc) == exp["sha256"], name
        off = 8 + 8 + 16 * ch
        for k, s in enumerate(exp["first_slices"]):
            assert enc[off + 8 * k: off + 8 * k + 8].hex() == s, (name, k)
        if len(sig) <= 2000:  # pure python is slow
            assert py.encode_all([int(v) for v in sig], ch, rate, n) == enc, name

    # Synthetic-generator pins (so the CUDA generator can be checked on the GPU box without the oracle drifting)
    synth = {}
    for kind in ("music", "noise"):
        for (seed, stream, ch, n) in [(1, 0, 2, 12345), (2, 17, 1, 5121), (3, 5, 8, 777)]:
            p = signals.synth_pcm(seed, stream, ch, n, kind)
            st, enc = ob.encode_all(p, ch, 44100, n)
            assert st == 0
            synth[f"{kind}:{seed}:{stream}:{ch}:{n}"] = {"pcm_sha256": sha(p.tobytes()), "qoa_sha256": sha(enc),
                                                         "qoa_len": len(enc)}
    g["synth"] = synth
    with open(os.path.join(ROOT, "tests/golden/golden.json"), "w") as f:
        json.dump(g, f, indent=1, sort_keys=True)
    print("wrote golden.json")


def parity_inputs():
    """name -> (pcm int16, channels, rate, samples): everything whose encoder output the oracle vouches for."""
    inputs = dict(signals.kat_signals())
    for kind in ("music", "noise"):
        for (seed, stream, ch, n) in [(1, 0, 2, 12345), (2, 17, 1, 5121), (3, 5, 8, 777)]:
            inputs[f"synth_{kind}_{seed}_{stream}_{ch}_{n}"] = (signals.synth_pcm(seed, stream, ch, n, kind), ch, 44100, n)
    for name, sig in signals.edge_signals().items():
        inputs[f"edge_{name}"] = (sig, 1, 8000, sig.size)
    return inputs


def export_inputs(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    for name, (sig, ch, rate, n) in parity_inputs().items():
        np.asarray(sig, dtype="<i2").tofile(os.path.join(out_dir, name + ".pcm"))
        with open(os.path.join(out_dir, name + ".desc"), "w") as f:
            f.write(f"{ch} {rate} {n}\n")
    fixture = open(os.path.join(ROOT, "tests/golden/julien_baker_sprained_ankle.qoa"), "rb").read()
    with open(os.path.join(out_dir, "fixture.qoa"), "wb") as f:
        f.write(fixture)
    print(f"exported {len(parity_inputs())} PCM inputs + the fixture to {out_dir}")


def check_against(crate_dir):
    """Compares what rust/golden-dump wrote (the reference crate's own output) with the oracle, byte for byte."""
    bad = 0
    for name, (sig, ch, rate, n) in parity_inputs().items():
        st, enc = ob.encode_all(sig, ch, rate, n)
        crate = open(os.path.join(crate_dir, name + ".qoa"), "rb").read()
        crate_dec = np.fromfile(os.path.join(crate_dir, name + ".dec.pcm"), dtype="<i2")
        ok = st == 0 and crate == bytes(enc) and np.array_equal(crate_dec, ob.decode_all(bytes(enc))[1])
        bad += not ok
        print(f"{'ok      ' if ok else 'MISMATCH'} {name}")
    fixture = open(os.path.join(ROOT, "tests/golden/julien_baker_sprained_ankle.qoa"), "rb").read()
    dec = np.fromfile(os.path.join(crate_dir, "fixture.decoded.pcm"), dtype="<i2")
    re_enc = open(os.path.join(crate_dir, "fixture.reencoded.qoa"), "rb").read()
    ok = (np.array_equal(dec, ob.decode_all(fixture)[1]) and sha(dec.tobytes()) == APPENDIX_B["fixture_decoded_pcm_sha256"]
          and sha(re_enc) == APPENDIX_B["fixture_reencode_sha256"])
    bad += not ok
    print(f"{'ok      ' if ok else 'MISMATCH'} fixture decode + re-encode")
    print("oracle PINNED to the crate" if not bad else f"{bad} mismatches: the oracle is NOT the crate")
    return 1 if bad else 0


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--export-inputs":
        export_inputs(sys.argv[2])
    elif len(sys.argv) == 3 and sys.argv[1] == "--check-against":
        sys.exit(check_against(sys.argv[2]))
    else:
        main()

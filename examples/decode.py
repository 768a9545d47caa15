#!/usr/bin/env python3
"""QOA -> 16-bit PCM WAV on the B200 engine: the counterpart of the reference's examples/decode.rs (:6-58), plus a batch
form that decodes any number of files in ONE engine call.

    decode.py <input.qoa> <output.wav>
    decode.py --out-dir DIR <input.qoa> [<input.qoa> ...]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(argv):
    if len(argv) >= 4 and argv[1] == "--out-dir":
        out_dir, inputs = argv[2], argv[3:]
        outputs = [os.path.join(out_dir, os.path.splitext(os.path.basename(p))[0] + ".wav") for p in inputs]
        os.makedirs(out_dir, exist_ok=True)
    elif len(argv) == 3:
        inputs, outputs = [argv[1]], [argv[2]]
    else:
        print(__doc__.strip(), file=sys.stderr)
        return 0
    from qoaudio_b200 import DecodeError, decode_all_batch, wavio
    blobs, good = [], []
    for k, p in enumerate(inputs):
        if not os.path.exists(p):
            print(f"Error: Input file '{p}' does not exist", file=sys.stderr)
            continue
        with open(p, "rb") as f:
            blobs.append(f.read())
        good.append(k)
    if not good:
        return 1
    print(f"Decoding {len(good)} QOA file(s)...")
    failed = len(inputs) - len(good)
    for k, res in zip(good, decode_all_batch(blobs)):
        if isinstance(res, DecodeError):
            print(f"Error: {inputs[k]}: {res}", file=sys.stderr)
            failed += 1
            continue
        per_channel = res.samples.size // res.num_channels
        wavio.write_wav(outputs[k], res.num_channels, res.sample_rate, res.samples)
        print(f"  {inputs[k]}: {res.num_channels} ch, {res.sample_rate} Hz, {per_channel} samples per channel, "
              f"{per_channel / res.sample_rate:.2f} s -> {outputs[k]} ({os.path.getsize(outputs[k])} bytes)")
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))

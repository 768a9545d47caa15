#!/usr/bin/env python3
"""WAV -> QOA on the B200 engine: the counterpart of the reference's examples/encode.rs (:9-95), plus a batch form that
encodes any number of files in ONE engine call (which is what the GPU is for).

    encode.py <input.wav> <output.qoa>
    encode.py --out-dir DIR <input.wav> [<input.wav> ...]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def load(path):
    """Returns (samples, QoaDesc) or an error string (the checks of examples/encode.rs:24-56)."""
    from qoaudio_b200 import QoaDesc, wavio
    if not os.path.exists(path):
        return f"Error: Input file '{path}' does not exist"
    try:
        channels, rate, bits, samples = wavio.read_wav(path)
    except (OSError, wavio.WavError) as e:
        return f"Error: cannot read '{path}': {e}"
    if samples is None:
        return f"Error: Only 16-bit PCM WAV files are supported. Found {bits} bits per sample"
    if channels == 0 or channels > 8:
        return f"Error: Unsupported number of channels: {channels}. Must be between 1 and 8"
    if rate == 0:
        return f"Error: Invalid sample rate: {rate}"
    per_channel = samples.size // channels
    return samples[:per_channel * channels], QoaDesc(channels, rate, per_channel)


def main(argv):
    if len(argv) >= 4 and argv[1] == "--out-dir":
        out_dir, inputs = argv[2], argv[3:]
        outputs = [os.path.join(out_dir, os.path.splitext(os.path.basename(p))[0] + ".qoa") for p in inputs]
        os.makedirs(out_dir, exist_ok=True)
    elif len(argv) == 3:
        inputs, outputs = [argv[1]], [argv[2]]
    else:
        print(__doc__.strip(), file=sys.stderr)
        print("\nThe input files must be 16-bit PCM WAV files.", file=sys.stderr)
        return 0
    from qoaudio_b200 import EncodeError, encode_all_batch
    loaded = [load(p) for p in inputs]
    for p, item in zip(inputs, loaded):
        if isinstance(item, str):
            print(item, file=sys.stderr)
            continue
        samples, desc = item
        print(f"{p}: {desc.channels} ch, {desc.sample_rate} Hz, {desc.samples} samples per channel, "
              f"{desc.samples / desc.sample_rate:.2f} s, {os.path.getsize(p)} bytes")
    good = [k for k, item in enumerate(loaded) if not isinstance(item, str)]
    if not good:
        return 1
    print(f"\nEncoding {len(good)} file(s) to QOA format...")
    results = encode_all_batch([loaded[k] for k in good])
    failed = len(loaded) - len(good)
    for k, res in zip(good, results):
        if isinstance(res, EncodeError):
            print(f"Error: {inputs[k]}: {res}", file=sys.stderr)
            failed += 1
            continue
        with open(outputs[k], "wb") as f:
            f.write(res)
        pcm_bytes = loaded[k][0].size * 2
        print(f"  {outputs[k]}: {len(res)} bytes, compression ratio {pcm_bytes / len(res):.2f}x")
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))

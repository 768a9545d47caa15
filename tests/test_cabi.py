"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/qoa_b200.h declares,
its pure-host helpers agree with the oracle, and compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from oracle import binding as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "qoa_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qoa_b200_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    import qoaudio_b200._native as N
    assert header_symbols() == sorted(N.SYMBOLS)


def test_rust_shim_declares_the_same_symbols():
    """rust/qoa-b200-sys (shipped as source: no rustc in this image) must bind exactly the header's entry points."""
    src = open(os.path.join(ROOT, "rust", "qoa-b200-sys", "src", "lib.rs")).read()
    assert sorted(set(re.findall(r"fn (qoa_b200_[a-z0-9_]+)", src))) == header_symbols()


def test_library_exports_every_declared_symbol():
    import qoaudio_b200._native as N
    L = N.lib()
    for name in header_symbols():
        assert hasattr(L, name), name


def test_size_formulas_match_oracle():
    import qoaudio_b200._native as N
    L = N.lib()
    for ch in range(1, 9):
        for n in (1, 19, 20, 21, 5119, 5120, 5121, 10259, 2394122, 14400000):
            assert L.qoa_b200_encoded_size(ch, n) == ob.encoded_size(ch, n)
    assert L.qoa_b200_frame_size(2, 5120) == 4136
    assert L.qoa_b200_frame_size(8, 5120) == 16520
    assert L.qoa_b200_frame_size(1, 1) == 32


def test_status_codes_shared_with_oracle():
    import qoaudio_b200._native as N
    for name in ("OK", "ENC_INVALID_CHANNELS", "ENC_INVALID_SAMPLE_RATE", "ENC_INVALID_SAMPLES", "ENC_IO_ERROR",
                 "DEC_NOT_QOA_FILE", "DEC_NO_SAMPLES", "DEC_INVALID_FRAME_HEADER", "DEC_INCOMPATIBLE_FRAME",
                 "DEC_IO_ERROR"):
        assert getattr(N, name) == getattr(ob, name)
    hdr = open(os.path.join(ROOT, "include", "qoa_b200.h")).read()
    for name, val in (("QOA_B200_DEC_IO_ERROR", 15), ("QOA_B200_ENC_IO_ERROR", 4), ("QOA_B200_DEC_NOT_QOA_FILE", 11)):
        assert re.search(rf"{name} = {val}\b", hdr)


def test_struct_layouts():
    import qoaudio_b200._native as N
    assert C.sizeof(N.Desc) == 12
    assert C.sizeof(N.StreamInfo) == 48
    assert C.sizeof(N.FrameHeaderC) == 12


def test_no_cpu_fallback():
    """Without a usable sm_100 device the engine refuses to start — it never computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import qoaudio_b200
    import qoaudio_b200._native as N
    with pytest.raises(N.NativeError):
        qoaudio_b200.Engine(0)
    with pytest.raises(N.NativeError):
        qoaudio_b200.encode_all([0] * 10, qoaudio_b200.QoaDesc(1, 44100, 10))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under qoaudio_b200/ may reference it."""
    pkg = os.path.join(ROOT, "qoaudio_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower(), os.path.join(dirpath, f)


def test_crate_surface_names():
    import qoaudio_b200 as q
    for name in ("encode_all", "decode_all", "open_and_decode_all", "QoaEncoder", "QoaDecoder", "QoaItem", "QoaDesc",
                 "FrameHeader", "ProcessingMode", "DecodedQoa", "EncodeError", "DecodeError", "QOA_FRAME_LEN",
                 "QOA_SLICE_LEN", "QOA_LMS_LEN", "QOA_HEADER_SIZE", "QOA_MAGIC", "QOA_SLICES_PER_FRAME",
                 "QOA_MAX_CHANNELS", "MAX_SLICES_PER_CHANNEL_PER_FRAME"):
        assert hasattr(q, name), name
    assert q.QOA_FRAME_LEN == 5120 and q.QOA_MAGIC == 0x716F6166
    for m in ("encode", "write_header", "encode_frame", "new"):
        assert hasattr(q.QoaEncoder, m)
    for m in ("new", "open", "new_streaming", "decode_frame", "mode", "current_frame_header", "total_duration",
              "into_inner"):
        assert hasattr(q.QoaDecoder, m)

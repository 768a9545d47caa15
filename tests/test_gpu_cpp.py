"""include/qoa_b200.hpp — the C++ mirror of the crate's API — exercised by tests/cpp/wrapper_check.cpp on the GPU; its
outputs are compared with the oracle here."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import binding as ob

CPP = os.path.join(ROOT, "tests", "cpp")


def test_header_compiles_standalone():
    """CPU: the header is self-contained and warning-free (no GPU needed to compile or link)."""
    subprocess.check_call(["make", "-s", "-C", CPP])
    assert os.path.exists(os.path.join(CPP, "_build", "wrapper_check"))


@pytest.mark.gpu
def test_cpp_wrapper_matches_oracle(tmp_path):
    subprocess.check_call(["make", "-s", "-C", CPP])
    r = subprocess.run([os.path.join(CPP, "_build", "wrapper_check"), str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == "OK", r.stdout + r.stderr
    for name, (ch, rate) in (("stereo", (2, 44100)), ("mono", (1, 8000))):
        pcm = np.fromfile(tmp_path / f"{name}.pcm", dtype=np.int16)
        qoa = (tmp_path / f"{name}.qoa").read_bytes()
        st, expect = ob.encode_all(pcm, ch, rate, pcm.size // ch)
        assert st == 0 and qoa == expect, name
        assert np.array_equal(np.fromfile(tmp_path / f"{name}.dec", dtype=np.int16), ob.decode_all(qoa)[1]), name
    res = dict(line.split("=") for line in (tmp_path / "results.txt").read_text().split())
    cut = (tmp_path / "stereo.qoa").read_bytes()[:int(res["trunc_len"])]
    info, opcm, _ = ob.decode_iter(cut)
    assert info["status"] == ob.DEC_IO_ERROR and int(res["trunc_samples"]) == info["samples"] == len(opcm)

"""The device header's permutation, compiled as HOST code (tests/hostcheck/keccak_header_host.cpp defines the CUDA
qualifiers and the two funnel-shift intrinsics away), against the oracle: pins the rho/pi/chi index tables and the
digest-only last round (keccak_last_round_digest) without a GPU.  The product never runs this way."""
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def test_device_header_permutation_on_host(oracle, tmp_path):
    exe = str(tmp_path / "kh")
    subprocess.run(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(HERE, "hostcheck", "keccak_header_host.cpp")], check=True)
    rng = np.random.default_rng(5)
    msgs = [bytes(rng.integers(0, 256, n, dtype=np.uint8)) for n in list(range(0, 300)) + [532, 543, 544, 545, 1000, 4096]]
    stdin = "\n".join(m.hex() if m else "-" for m in msgs) + "\n"
    out = subprocess.run([exe], input=stdin, capture_output=True, text=True, check=True).stdout.split("\n")
    for m, line in zip(msgs, out):
        full, digest_only = line.split()
        want = oracle.keccak256(m).hex()
        assert full == want, len(m)
        assert digest_only == want, len(m)

"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and refuses to work without a CUDA device (no CPU fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "phant_gpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(phant_gpu_[a-z0-9_]+)\s*\(", hdr)))


def test_header_declares_the_survey_entry_points():
    syms = declared_symbols()
    for need in ["phant_gpu_abi_version", "phant_gpu_create", "phant_gpu_destroy", "phant_gpu_strerror",
                 "phant_gpu_keccak256_batch", "phant_gpu_mpt_root", "phant_gpu_state_root", "phant_gpu_verify_proofs",
                 "phant_gpu_trie_open", "phant_gpu_trie_update", "phant_gpu_trie_close"]:
        assert need in syms


def test_library_exports_every_declared_symbol():
    from phant_b200 import gpu
    import ctypes
    lib = ctypes.CDLL(gpu.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    assert set(declared_symbols()) == set(gpu.EXPORTS)
    assert gpu.abi_version() == 2


def test_strerror_covers_all_codes():
    from phant_b200 import gpu
    L = gpu._lib()
    for code in range(0, -7, -1):
        assert L.phant_gpu_strerror(code) not in (None, b"unknown error")
    assert L.phant_gpu_strerror(-99) == b"unknown error"


def test_no_silent_cpu_fallback():
    """Without a CUDA device the product must fail loudly, not route through a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from phant_b200 import gpu
    with pytest.raises(gpu.PhantGpuError) as e:
        gpu.Context(0)
    assert e.value.code == -2


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    walks = [w for d in ("phant_b200", "host", "include", "zig") for w in os.walk(os.path.join(ROOT, d))]
    for dirpath, _, files in walks:
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_lib" not in txt and "liboracle" not in txt and "oracle/" not in txt.replace("oracle/synth.c", "").replace("oracle/verify.c", ""), os.path.join(dirpath, f)


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """gcc -std=c99 -pedantic on the header, link against the library, run: NO_DEVICE here, a real hash on a GPU box."""
    import subprocess
    lib = os.path.join(ROOT, "phant_b200", "lib")
    exe = str(tmp_path / "abi_c99_check")
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-o", exe, os.path.join(ROOT, "tests", "abi_c99_check.c"), f"-L{lib}",
                    "-lphantgpu", f"-Wl,-rpath,{lib}"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)


def test_cpp_host_mirror_compiles_and_links_without_a_gpu(tmp_path):
    """host/phant_host.hpp + host/host_test.cpp build against the header and link against the library on a machine with no
    GPU (running them is the -m gpu test tests/test_gpu_host_cpp.py)"""
    import subprocess
    from phant_b200 import gpu
    lib = os.path.dirname(gpu.LIB_PATH)
    subprocess.run(["g++", "-std=c++17", "-O0", "-Wall", "-Werror", "-Wno-unused-function", "-o", str(tmp_path / "host_test"),
                    os.path.join(ROOT, "host", "host_test.cpp"), f"-L{lib}", "-lphantgpu", f"-Wl,-rpath,{lib}"], check=True)

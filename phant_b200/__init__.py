"""phant_b200 -- phant's trie/hash hot path on NVIDIA B200.

The product is the CUDA library `lib/libphantgpu.so` behind the C ABI of `include/phant_gpu.h`.
This package is the thin Python host layer used by the tests and the benchmark: it mirrors the names
of the reference functions the library stands behind (src/crypto/hasher.zig, src/mpt/mpt.zig,
src/blockchain/blockchain.zig:209-235, the StateDB.root() / witness hooks) and does no arithmetic of
its own.  There is no CPU fallback: importing `phant_b200.gpu` without the built library, or creating
a context without a CUDA device, raises.
"""
from . import gpu  # noqa: F401
from .gpu import Context, PhantGpuError  # noqa: F401

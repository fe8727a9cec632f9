#!/usr/bin/env python3
"""sha256 of each kernel's SASS in libphantgpu.so (addresses and encodings stripped, anonymous-namespace hashes folded):
a cheap way to tell whether a source change altered the machine code that was verified on the GPU.
  python tools/sass_fingerprint.py > profiles/sass_fingerprint_<tag>.txt
"""
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "phant_b200", "lib", "libphantgpu.so")
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    out = re.sub(r"_GLOBAL__N__[0-9a-f]+_[0-9]+_[a-z_0-9]+_cu_[0-9a-f]+", "ANON", out)
    kernels, cur = {}, None
    for line in out.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = hashlib.sha256()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?)\s*/\* 0x[0-9a-f]+ \*/", line)
        if m and cur:
            kernels[cur].update(m.group(1).encode() + b"\n")
    for name in sorted(kernels):
        print(kernels[name].hexdigest()[:16], name[:120])


if __name__ == "__main__":
    main()

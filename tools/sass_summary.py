"""SASS mnemonic counts per kernel of the in-tree libsyldet_cuda.so (cuobjdump -sass | c++filt): what profiles/*_sass_summary.txt holds.

    python tools/sass_summary.py > profiles/rNN_sass_summary.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "syllable-detector-swift_b200", "libsyldet_cuda.so")
KEYS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UBLKCP", "UBLKPF", "LDTM", "STTM", "UTCATOMSWS", "USETMAXREG", "SYNCS", "FFMA", "FFMA2", "FADD2", "FMUL2",
        "FHFMA", "MUFU", "SHFL", "LDS", "STS", "LDG", "STG"]

sass = subprocess.run("cuobjdump -sass %s | c++filt" % LIB, shell=True, capture_output=True, text=True).stdout
kernels, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (.*)", line)
    if m:
        name = m.group(1)
        name = re.sub(r"syldet::\(anonymous namespace\)::|syldet::|void ", "", name)
        name = re.sub(r"\(.*", "", name)
        cur = kernels.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        cur["total"] += 1
        cur[m.group(1)] += 1
print("SASS mnemonic counts per kernel of libsyldet_cuda.so (cuobjdump -sass; tools/sass_summary.py). UTCHMMA = tcgen05.mma (kind::tf32 / kind::f16),")
print("UTMALDG = cp.async.bulk.tensor, UBLKCP = cp.async.bulk, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops,")
print("FHFMA = mixed-precision fma.rn.f32.f16, FADD2 / FMUL2 / FFMA2 = packed fp32 (two results per instruction).\n")
for name, c in kernels.items():
    print("%-64s total %6d  %s" % (name[:64], c["total"], " ".join("%s=%d" % (k, c[k]) for k in KEYS if c[k])))

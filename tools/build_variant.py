"""Build an A/B variant of the library: every translation unit as in nabo_b200.build, with extra -D switches,
into tools/_variants/lib_<name>.so (select it at run time with NABO_B200_LIB=<path>).  Development probe."""
import os, subprocess, sys
sys.path.insert(0, ".")
from nabo_b200 import build as B
name, defs = sys.argv[1], sys.argv[2:]
out = os.path.join("tools", "_variants")
os.makedirs(os.path.join(out, name), exist_ok=True)
objs = []
procs = []
for src in B.sources():
    obj = os.path.join(out, name, os.path.basename(src)[:-3] + ".o")
    objs.append(obj)
    procs.append(subprocess.Popen(["nvcc", *B.ARCH_FLAGS, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
                                   "--expt-relaxed-constexpr", *defs, "-c", src, "-o", obj]))
assert all(p.wait() == 0 for p in procs)
lib = os.path.join(out, "lib_%s.so" % name)
subprocess.check_call(["nvcc", *B.ARCH_FLAGS, "-shared", "-o", lib, *objs, "-lcuda"])
print(lib)

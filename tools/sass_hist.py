"""SASS opcode histograms of the library's hot kernels (profiles/r2_sass_*.txt):
python tools/sass_hist.py <name regex> "<title>" > profiles/r2_sass_<what>.txt"""
import collections, re, subprocess, sys
pat, title = re.compile(sys.argv[1]), sys.argv[2]
lib = "nabo_b200/libnabo_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
print("SASS opcode histogram - %s" % title)
print("source: cuobjdump -sass %s (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo), this commit\n" % lib)
FAMILIES = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTCCP", "UBLKCP", "UTMALDG", "SYNCS", "DMMA", "HMMA", "FMNMX3",
            "REDUX", "ATOMS", "ATOMG", "RED", "VOTE", "SHFL", "LOP3", "LDS", "STS", "LDG", "STG", "DADD", "DMUL", "DFMA", "MUFU", "BAR"]
cur, body = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1) if pat.search(m.group(1)) else None
        if cur:
            body[cur] = []
        continue
    if cur:
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            body[cur].append(m.group(1))
for name, ops in body.items():
    base = collections.Counter(o.split(".")[0] for o in ops)
    print("== %s  (%d instructions)" % (name, len(ops)))
    ev = ["%s x%d" % (f, base[f]) for f in FAMILIES if base.get(f)]
    print("   Blackwell / pipeline evidence: " + ", ".join(ev))
    full = sorted({o for o in ops if o.split(".")[0] in ("UTCHMMA", "LDTM", "UBLKCP", "SYNCS", "DMMA", "REDUX", "UTCBAR", "FMNMX3") and "." in o})
    if full:
        print("   full mnemonics of those families: " + ", ".join(full))
    print("   opcode histogram (base mnemonic: count): " + ", ".join("%s:%d" % kv for kv in base.most_common(40)))
    print()

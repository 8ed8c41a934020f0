"""Per-kernel totals and shares from an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X):
python tools/launch_shares.py X.csv "header text" > profiles/..._launch_shares.txt"""
import collections, csv, re, sys
path, header = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
    name = re.sub(r"\(.*", "", r[ik])
    name = re.sub(r"\(int\)|\(bool\)", "", name)[:60]
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
if header:
    print(header)
print("source: ncu --metrics gpu__time_duration.sum --clock-control none --csv (%s); cold-cache, serialised: compare SHARES\n" % path)
print("%-60s %8s %12s %8s %12s" % ("kernel", "launches", "total us", "share", "us / launch"))
for k, v in tot.most_common():
    if v / total < 2e-4:
        continue
    print("%-60s %8d %12.1f %7.2f%% %12.1f" % (k, cnt[k], v, 100 * v / total, v / cnt[k]))

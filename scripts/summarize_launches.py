"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total us, share."""
import collections, csv, sys
src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(src)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) < 15:
        continue
    name = r[4].split('(')[0].replace('void ', '')[:60]
    key = name + (" grid" + r[8].replace(", ", "x") if 'gemm' in name else "")
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += float(r[14])
tot = sum(v[1] for v in agg.values())
out = ["# " + title, "# ncu per-launch times are cold-cache and serialised: compare SHARES, not absolutes",
       "# kernel, launches, total_us, share"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("%s, %d, %.1f, %.4f" % (k, v[0], v[1] / 1e3, v[1] / tot))
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out[:12]))

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total us, share.
usage: summarize_launches.py launches.csv out.csv "title" [step]
With `step`, only the launches of ONE training step are kept: those between the step-th and the (step+1)-th
rng_tick_kernel (the first kernel of every step)."""
import collections, csv, sys
src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
step = int(sys.argv[4]) if len(sys.argv) > 4 else None
rows = list(csv.reader(open(src)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
body = [r for r in rows[hdr + 1:] if len(r) >= 15]
if step is not None:
    ticks = [i for i, r in enumerate(body) if "rng_tick_kernel" in r[4]]
    body = body[ticks[step]:ticks[step + 1]]
agg = collections.OrderedDict()
for r in body:
    name = r[4].split('(')[0].replace('void ', '')[:60]
    key = name + (" grid" + r[8].replace(", ", "x") if 'gemm' in name else "")
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += float(r[14])
tot = sum(v[1] for v in agg.values())
fam = collections.OrderedDict()
for k, v in agg.items():
    f = k.split(" grid")[0].split("<")[0]
    fam[f] = fam.get(f, 0.0) + v[1]
out = ["# " + title, "# ncu per-launch times are cold-cache and serialised: compare SHARES, not absolutes",
       "# %d launches, %.1f us in total" % (sum(v[0] for v in agg.values()), tot / 1e3),
       "# by kernel family: " + "; ".join("%s %.3f" % (k, v / tot) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])[:8]),
       "# kernel, launches, total_us, share"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("%s, %d, %.1f, %.4f" % (k, v[0], v[1] / 1e3, v[1] / tot))
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out[:14]))

"""Summarises an ncu launch list (--metrics gpu__time_duration.sum --csv) for the last complete denoiser forward in it:
per kernel family launches, total and mean duration, share of the forward.
python tools/launch_list_summary.py file.csv [first-kernel-substring last-kernel-substring]   (default: PoseNet's
pack_tokens .. unpack_tokens; TrajNet: pack_rows unpack_rows)"""
import collections, csv, sys

rows = list(csv.reader(open(sys.argv[1])))
i0 = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[i0], rows[i0 + 1:]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
names, durs = [], []
for r in data:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(',', ''))
    names.append(r[kn])
    durs.append(v / 1000 if r[mu] == 'ns' else v)
first, last = (sys.argv[2], sys.argv[3]) if len(sys.argv) > 3 else ('pack_tokens', 'unpack_tokens')
un = [i for i, n in enumerate(names) if last in n]
pk = [i for i, n in enumerate(names) if first in n and last not in n]
b = un[-1]
a = max(i for i in pk if i < b)
nxt = [i for i in pk if i > b]
agg = collections.OrderedDict()
for n, d in zip(names[a:b + 1], durs[a:b + 1]):
    key = n.split('(')[0].split('::')[-1][:60]
    agg.setdefault(key, [0, 0.0])
    agg[key][0] += 1
    agg[key][1] += d
tot = sum(durs[a:b + 1])
print(f"one forward = launches {a}..{b} ({b + 1 - a} kernels), sum of durations {tot:.1f} us (cold caches, serialised)")
for k, (c, d) in agg.items():
    print(f"  {k:62s} x{c:3d} {d:8.1f} us {100 * d / tot:5.1f}%  mean {d / c:6.1f} us")
if b + 1 < len(names):
    print("between forwards (sampler step):")
    for n, d in list(zip(names, durs))[b + 1:(nxt[0] if nxt else b + 6)]:
        print(f"  {d:7.1f} us  {n[:100]}")

#!/usr/bin/env python
"""Aggregate an ncu source-page CSV of the solve kernel by device function and stall reason.
   python scripts/ncu_by_function.py <source.csv> <nvdisasm -c output> <kernel mangled name> <evals>"""
import bisect, collections, csv, re, sys
src, sass, kname, evals = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
lines = open(sass).read().split('\n')
start = [i for i, l in enumerate(lines) if l.startswith('.text.' + kname + ':')][0]
funcs, cur = [(0, 'kernel main')], None
for l in lines[start + 1:]:
    if l.startswith('.text.') or l.startswith('.section'):
        break
    m = re.match(r'^(\S+):\s*$', l)
    if m:
        cur = m.group(1); continue
    m = re.match(r'^\s+/\*([0-9a-f]{4,6})\*/', l)
    if m and cur:
        if cur.startswith('$'):
            name = cur.split('$')[-1]
            name = re.sub(r'^_ZN3dib\d+', '', name)
            funcs.append((int(m.group(1), 16), name[:28]))
        cur = None
rows = list(csv.reader(open(src)))
hdr = rows[1]
ia, isamp, iex = hdr.index('Address'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stalls = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for r in rows[2:]:
    try:
        data.append((int(r[ia], 16), int(r[isamp]), int(r[iex]), r))
    except ValueError:
        pass
base = data[0][0]
addrs = [a for a, _ in funcs]
agg = collections.OrderedDict((n, [0, 0, collections.Counter()]) for _, n in funcs)
tot = sum(d[1] for d in data); totex = sum(d[2] for d in data)
allst = collections.Counter()
for a, s, e, r in data:
    f = funcs[bisect.bisect_right(addrs, a - base) - 1][1]
    agg[f][0] += s; agg[f][1] += e
    for i, h in stalls:
        if r[i] not in ('', '0'):
            agg[f][2][h] += int(r[i]); allst[h] += int(r[i])
print("total samples %d, executed warp-instr %.3e (%.0f per evaluation)" % (tot, totex, totex / evals))
print("stall mix: " + ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in allst.most_common(9)))
for n, (s, e, st) in agg.items():
    if e:
        print("%-30s samples %5.2f%%  executed/eval %8.1f  top stalls: %s" % (
            n, 100 * s / tot, e / evals, ", ".join("%s %.0f%%" % (k, 100 * v / max(s, 1)) for k, v in st.most_common(3))))

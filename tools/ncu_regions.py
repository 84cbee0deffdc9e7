"""Aggregate an ncu source-page CSV by source line / file using nvdisasm -g line info.
usage: python tools/ncu_regions.py src.csv dis.txt"""
import re, csv, collections, sys
addr2line = {}
cur = None
for ln in open(sys.argv[2]):
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S+)', ln)
    if m: addr2line[int(m.group(1), 16)] = (cur, m.group(2))
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ia = hdr.index("Address"); ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
byline = collections.Counter(); bysamp = collections.Counter(); byop = collections.Counter(); stalls = collections.Counter()
tot = tots = 0; base = None
for r in rows[2:]:
    try: a = int(r[ia], 16)
    except Exception: continue
    if base is None: base = a
    e = int(r[ie] or 0); s = int(r[isamp] or 0)
    key, op = addr2line.get(a - base, (None, "?"))
    byline[key] += e; bysamp[key] += s; tot += e; tots += s
    byop[op.split('.')[0]] += e
    for i in stall_cols:
        try: stalls[hdr[i]] += int(r[i] or 0)
        except ValueError: pass
print("total warp-instructions", tot, "samples", tots)
byfile = collections.Counter(); sfile = collections.Counter()
for k, v in byline.items(): byfile[k[0] if k else "?"] += v
for k, v in bysamp.items(): sfile[k[0] if k else "?"] += v
for k, v in byfile.most_common(8): print("%-24s inst %5.1f%%  samples %5.1f%%" % (k, 100 * v / tot, 100 * sfile[k] / tots))
print("top lines (by samples):")
for k, v in bysamp.most_common(30): print("  ", k, "samples %.1f%%" % (100 * v / tots), "inst %.1f%%" % (100 * byline[k] / tot))
print("opcodes:", [(k, "%.1f%%" % (100 * v / tot)) for k, v in byop.most_common(14)])
print("stalls:", [(k, "%.1f%%" % (100 * v / max(1, sum(stalls.values())))) for k, v in stalls.most_common(8)])

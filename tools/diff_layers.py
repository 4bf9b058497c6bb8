"""Row-by-row difference of two tools/profile_layers.py tables: python tools/diff_layers.py OLD NEW"""
import sys


def load(f):
    sec, d = None, {}
    for l in open(f):
        if l.startswith("=="):
            sec = l.split(":")[0][3:]
            continue
        p = l.split()
        if len(p) >= 13 and p[1].isdigit():
            d[(sec,) + tuple(p[:8])] = float(p[9])
    return d


a, b = load(sys.argv[1]), load(sys.argv[2])
rows = sorted(((b[k] - a[k], k, a[k], b[k]) for k in a if k in b))
print(f"sum old {sum(a.values()):.2f} ms, new {sum(b.values()):.2f} ms")
for r in rows[:10] + rows[-8:]:
    print("%+.2f ms" % r[0], " ".join(r[1]), r[2], "->", r[3])

"""Per-step-kind cycle breakdown of a walk trace (HB2_WALK_TRACE output)."""
import sys
rows = [l.split() for l in open(sys.argv[1]) if not l.startswith('#')]
L = int(sys.argv[2]) if len(sys.argv) > 2 else 200
cat = {}
prev_end = None
for r in rows:
    enc = int(r[1], 16); fl = int(r[2], 16)
    q = [None] + [int(x) for x in r[3:]]
    child = enc & 0x0fffffff
    kind = 'leaf' if child < L else ('wait' if enc & (1 << 30) else ('chain' if enc & (1 << 29) else 'int'))
    if fl & (1 << 29): kind += '+last'
    d = cat.setdefault(kind, {'n': 0, 'pre': 0, 'body': 0, 'conv': 0, 'bar2': 0, 'mma': 0, 'epi': 0, 'gap': 0})
    d['n'] += 1
    d['pre'] += q[2] - q[1]
    d['body'] += q[7] - q[2]
    if not kind.startswith('leaf'):
        d['conv'] += q[3] - q[2]; d['bar2'] += q[4] - q[3]; d['mma'] += q[6] - q[4]; d['epi'] += q[7] - q[6]
    if prev_end is not None: d['gap'] += q[1] - prev_end
    prev_end = q[7]
for k, d in sorted(cat.items()):
    n = d['n']; print(f"{k:12s} n={n:3d} total={sum(v for a, v in d.items() if a in ('pre','body','gap')):8d}", {a: round(b / n) for a, b in d.items() if a != 'n'})
print('lane total', rows[-1][9])

"""Groups consecutive SASS instructions with the same executed count (ncu source-page CSV) -> basic-block level dynamic profile."""
import csv, sys
path = sys.argv[1]; unit = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(open(path)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
I = lambda r, k: int(r[ix[k]] or 0)
groups = []
for n, r in enumerate(data):
    ie = I(r, 'Instructions Executed')
    if groups and groups[-1]['ie'] == ie:
        g = groups[-1]; g['n'] += 1; g['samp'] += I(r, '# Samples'); g['bar'] += I(r, 'stall_barrier'); g['ops'].append(r[1].split()[0] if not r[1].strip().startswith('@') else r[1].split()[1])
    else:
        groups.append({'start': n, 'ie': ie, 'n': 1, 'samp': I(r, '# Samples'), 'bar': I(r, 'stall_barrier'), 'ops': [r[1].split()[0] if not r[1].strip().startswith('@') else r[1].split()[1]], 'first': r[1].strip()})
tot = sum(g['ie'] * g['n'] for g in groups)
cum = 0
for g in groups:
    if g['ie'] == 0: continue
    c = g['ie'] * g['n']; cum += c
    from collections import Counter
    ops = Counter(o.split('.')[0] for o in g['ops']).most_common(4)
    print(f"{g['start']:5d} n={g['n']:4d} exec/unit={g['ie']/unit:7.3f} share={100*c/tot:5.2f}% cum={100*cum/tot:5.1f}% samp={g['samp']:4d} bar={g['bar']:4d} {ops} | {g['first'][:44]}")

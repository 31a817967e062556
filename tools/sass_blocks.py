"""Per-block summary of an `ncu --page source --csv --print-source sass` export: executed warp instructions and stall samples."""
import csv, sys
path = sys.argv[1]; blk = int(sys.argv[2]) if len(sys.argv) > 2 else 50
rows = list(csv.reader(open(path)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
I = lambda r, k: int(r[ix[k]] or 0)
tot_inst = sum(I(r, 'Instructions Executed') for r in data)
tot_samp = sum(I(r, '# Samples') for r in data)
print('total warp inst', tot_inst, 'samples', tot_samp, 'n sass', len(data))
for b in range(0, len(data), blk):
    seg = data[b:b + blk]
    ie = sum(I(r, 'Instructions Executed') for r in seg)
    sm = sum(I(r, '# Samples') for r in seg)
    bar = sum(I(r, 'stall_barrier') for r in seg)
    print(f"{b:5d} inst={ie:9d} {100*ie/tot_inst:5.1f}%  samples={sm:6d} {100*sm/tot_samp:5.1f}% barrier={bar:5d}  first: {seg[0][1].strip()[:50]}")

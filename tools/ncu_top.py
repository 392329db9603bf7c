"""Top stalled SASS instructions from `ncu --page source --csv --print-source sass` output."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
h = rows[hi[0]]
end = hi[1] - 1 if len(hi) > 1 else len(rows)
body = [r for r in rows[hi[0] + 1:end] if len(r) == len(h)]
S = h.index('# Samples'); SRC = h.index('Source'); IE = h.index('Instructions Executed')
stall_cols = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
tot = sum(int(r[S]) for r in body)
print('total samples', tot, 'instructions', len(body))
for r in sorted(body, key=lambda r: -int(r[S]))[:topn]:
    st = sorted(((int(r[i]), h[i]) for i in stall_cols), reverse=True)[:2]
    print(f'{int(r[S]):7d} {100*int(r[S])/tot:5.1f}%  exec={r[IE]:>9}  {r[SRC].strip()[:70]:70s} {st}')

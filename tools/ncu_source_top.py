"""Top source lines by warp-stall samples from `ncu -i rep --page source --print-source cuda,sass --csv` (per function)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 14
seen = set()
i = 0
fn = None
agg = None
out = collections.OrderedDict()
for r in rows:
    if r and r[0] == "Function Name":
        fn = r[1][:60]
        continue
    if r and r[0] == "File Path":
        continue
    if r and r[0] == "Line No":
        continue
    if fn is None or len(r) < 6:
        continue
    if r[0] and r[0].isdigit():                      # a source line summary row
        try:
            v = float(r[4])
        except ValueError:
            continue
        out.setdefault(fn, collections.Counter())[(int(r[0]), r[1][:120])] += v
for fn, c in out.items():
    if want and want not in fn:
        continue
    if fn in seen:
        continue
    seen.add(fn)
    tot = sum(c.values())
    print(f"== {fn}  samples {tot:.0f}")
    for (ln, src), v in c.most_common(topn):
        print(f"  {v / tot * 100:5.1f}%  L{ln:<5d} {src}")

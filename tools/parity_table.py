"""Condense a PDE_PARITY_LOG (tests/conftest.py) into the table kept under profiles/: per test, the largest
measured error / bar ratio with its error and bar.  python tools/parity_table.py gpurun_out/parity.tsv > profiles/x.txt"""
import collections
import sys

rows = collections.OrderedDict()
n = 0
for line in open(sys.argv[1]):
    test, kind, what, err, tol = line.rstrip("\n").split("\t")
    err, tol = float(err), float(tol)
    n += 1
    key = (test, kind)
    if key not in rows or err / tol > rows[key][0] / rows[key][1]:
        rows[key] = (err, tol, what)
print(f"{n} parity comparisons in {len(set(k[0] for k in rows))} tests; per test and kind the comparison closest to its bar")
print(f"{'test':100s} {'kind':5s} {'error':>10s} {'bar':>10s} {'err/bar':>8s}  quantity")
worst = 0.0
for (test, kind), (err, tol, what) in rows.items():
    worst = max(worst, err / tol)
    print(f"{test[-100:]:100s} {kind:5s} {err:10.2e} {tol:10.2e} {err / tol:8.3f}  {what}")
print(f"largest error / bar over everything: {worst:.3f}")

"""Instruction mix / hottest SASS lines / stall reasons from an .ncu-rep source page (read here, no GPU)."""
import collections, csv, re, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS, iSamp, iEx = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
op, ops, st, data = collections.Counter(), collections.Counter(), collections.Counter(), []
tot_ex = tot_s = 0
for n, r in enumerate(rows[2:]):
    try:
        ex, s = int(r[iEx]), int(r[iSamp])
    except Exception:
        continue
    tot_ex += ex; tot_s += s
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iS])
    k = m.group(2).split('.')[0] if m else '?'
    op[k] += ex; ops[k] += s
    data.append((s, ex, n, r[iS].strip()))
    for i in stall_cols:
        try: st[hdr[i]] += int(r[i])
        except Exception: pass
print("warp instructions", tot_ex, "samples", tot_s)
for k, v in op.most_common(24):
    print(f"{k:12s} inst {v / tot_ex * 100:5.1f}%  samples {ops[k] / tot_s * 100:5.1f}%")
print({k: v for k, v in st.most_common(8)})
for s, e, n, src in sorted(data, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{s / tot_s * 100:5.2f}% ex={e:>9d} #{n:5d} {src[:100]}")

"""Instruction budget of tc_kernel from an ncu capture: executed warp instructions and stall samples per code
section, obtained by aligning the capture's SASS (ncu --page source) with the line / inlining information of the
same build (`nvdisasm --print-line-info-inline` on the cubin extracted from libpde_b200.so).  Read here, no GPU.

    python tools/sass_budget.py gpurun_out/prof.ncu-rep <cubin> <mangled kernel name> [points]
"""
import collections
import csv
import re
import subprocess
import sys

rep, cubin, kname = sys.argv[1], sys.argv[2], sys.argv[3]
points = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 22

dis = subprocess.run(["nvdisasm", "--print-line-info-inline", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.strip().startswith(".section") and ".text." + kname in l)
end = next((i for i in range(start + 1, len(dis)) if dis[i].strip().startswith(".section")), len(dis))
ins, chain, fresh = [], [], True
for l in dis[start:end]:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh:
            chain, fresh = [], False
        chain.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip(), tuple(chain)))
        fresh = True

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iA, iS, iSamp, iEx = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = rows[2:]
base = int(data[0][iA], 16)
assert len(data) == len(ins), (len(data), len(ins))


def opcode(t):
    return re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]


for (off, txt, _), r in zip(ins, data):
    assert int(r[iA], 16) - base == off and opcode(txt) == opcode(r[iS].strip()), (off, txt, r[iS])

# ---- sections: line ranges of pde_tc.cu (root = outermost frame) and leaf helpers
src = open("neural-network-based-pde-solver_b200/csrc/pde_tc.cu").read().splitlines()


def find(pat, after=0):
    return next(i + 1 for i, l in enumerate(src) if i + 1 > after and pat in l)


L_issuer, L_epi = find("if (warp >= NEPI) {"), find("// epilogue warps")
L_flush, L_loadx = find("auto flush_grads"), find("auto load_x")
L_fwd, L_out = find("auto fwd_layer"), find("// ================= output layer + envelope")
L_bwd, L_res = find("auto bwd_layer"), find("// ================= per-CTA results")
L_refill = find("auto refill = [&]", L_bwd)   # the rebuild of A_{l-1} from its stash
L_refill_end = find("        };", L_refill)
core = open("neural-network-based-pde-solver_b200/csrc/pde_tc_core.cuh").read().splitlines()
L_split = next(i + 1 for i, l in enumerate(core) if "void split2" in l)
L_sincos1, L_pk = find("void sincos_cw("), find("typedef unsigned long long f32x2;")
L_sincos2, L_acteval = find("void sincos_cw2"), find("void act_eval")


def outer(chain):
    for f, ln in reversed(chain):
        if f == "pde_tc.cu" and ln >= L_issuer:
            return ln
    return chain[-1][1] if chain else 0


def section(chain, txt):
    lines = [ln for f, ln in chain if f == "pde_tc.cu"]
    corel = [ln for f, ln in chain if f == "pde_tc_core.cuh"]
    root = outer(chain)
    if L_issuer <= root < L_epi:
        where = "issuer"
    elif any(L_flush <= ln < L_loadx for ln in lines):
        where = "flush_grads"
    elif any(L_fwd <= ln < L_out for ln in lines):
        where = "forward"
    elif any(L_refill <= ln < L_refill_end for ln in lines):
        where = "reverse/rebuild A"
    elif any(L_bwd <= ln < L_res for ln in lines):
        where = "reverse"
    elif any(L_out <= ln < L_bwd for ln in lines):
        where = "program"
    elif root >= L_res:
        where = "results"
    else:
        where = "tile setup"
    op = opcode(txt)
    if any(ln >= L_split for ln in corel):
        what = "hi/lo split"
    elif any(L_sincos2 <= ln < L_acteval + 25 or L_sincos1 <= ln < L_pk for ln in lines):
        what = "sin/cos"
    elif op in ("STSM",):
        what = "stmatrix"
    elif op in ("LDTM",) or any(136 <= ln <= 151 for ln in corel):
        what = "tmem ld"
    elif op in ("STG", "LDG"):
        what = "stash / global"
    elif op in ("SYNCS", "MEMBAR", "FENCE", "BAR", "WARPSYNC", "NANOSLEEP") or any(40 <= ln <= 69 for ln in corel):
        what = "barriers / fences"
    elif op in ("UTCHMMA", "UTCBAR", "R2UR", "UMOV", "UIADD3", "ELECT", "PLOP3") and where == "issuer":
        what = "mma issue"
    else:
        what = "chain rule / other"
    return where, what


ex = collections.Counter()
smp = collections.Counter()
tot_ex = tot_s = 0
for (off, txt, chain), r in zip(ins, data):
    try:
        e, s = int(r[iEx]), int(r[iSamp])
    except ValueError:
        continue
    k = section(chain, txt)
    ex[k] += e; smp[k] += s
    tot_ex += e; tot_s += s

units = points * 64 * 4        # point x hidden unit x layer
print(f"warp instructions {tot_ex:,}  ({tot_ex * 32 / units:.1f} thread instructions per point.unit.layer), samples {tot_s:,}")
print(f"{'section':22s} {'part':22s} {'inst %':>7s} {'thr-inst/p.u.l':>15s} {'samples %':>10s}")
wh = collections.Counter(); ws = collections.Counter()
for (where, what), e in sorted(ex.items(), key=lambda kv: -kv[1]):
    wh[where] += e; ws[where] += smp[(where, what)]
    if e / tot_ex >= 0.003 or smp[(where, what)] / tot_s >= 0.003:
        print(f"{where:22s} {what:22s} {e / tot_ex * 100:7.2f} {e * 32 / units:15.2f} {smp[(where, what)] / tot_s * 100:10.2f}")
print()
for where, e in wh.most_common():
    print(f"{where:22s} {'(all)':22s} {e / tot_ex * 100:7.2f} {e * 32 / units:15.2f} {ws[where] / tot_s * 100:10.2f}")
wt = collections.Counter(); wts = collections.Counter()
for (where, what), e in ex.items():
    if where != "issuer":
        wt[what] += e; wts[what] += smp[(where, what)]
print()
for what, e in wt.most_common():
    print(f"{'epilogue (all)':22s} {what:22s} {e / tot_ex * 100:7.2f} {e * 32 / units:15.2f} {wts[what] / tot_s * 100:10.2f}")

if len(sys.argv) > 5:      # top source lines of one section: python tools/sass_budget.py rep cubin kernel points "tile setup"
    want = sys.argv[5]
    byline = collections.Counter()
    for (off, txt, chain), r in zip(ins, data):
        try:
            e = int(r[iEx])
        except ValueError:
            continue
        if section(chain, txt)[0] == want:
            byline[" <- ".join(f"{f.replace('pde_tc', 'tc')}:{ln}" for f, ln in chain[:3])] += e
    print(f"\ntop lines of '{want}'")
    for k, e in byline.most_common(25):
        print(f"{e / tot_ex * 100:6.2f}%  {k}")

# ---- stall reasons of the epilogue warps (issuer and the idle warps parked at the final barrier excluded)
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
st = collections.Counter()
st_by = collections.defaultdict(collections.Counter)
for (off, txt, chain), r in zip(ins, data):
    where = section(chain, txt)[0]
    if where in ("issuer", "results"):
        continue
    for i in stall_cols:
        try:
            st[hdr[i]] += int(r[i]); st_by[where][hdr[i]] += int(r[i])
        except ValueError:
            pass
tot = sum(st.values())
print("\nepilogue warp samples by stall reason:", {k: f"{v / tot * 100:.1f}%" for k, v in st.most_common(9)})
for where, c in st_by.items():
    t = sum(c.values())
    print(f"  {where:20s} {t / tot * 100:5.1f}% of samples: ", {k.replace('stall_', ''): f"{v / t * 100:.0f}%" for k, v in c.most_common(6)})

if len(sys.argv) > 6:   # top instructions by one stall reason, e.g. ... 4194304 reverse stall_long_sb
    col = hdr.index(sys.argv[6])
    rowsx = []
    for (off, txt, chain), r in zip(ins, data):
        if section(chain, txt)[0] in ("issuer", "results"):
            continue
        try:
            rowsx.append((int(r[col]), off, txt, " <- ".join(f"{f.replace('pde_tc', 'tc')}:{ln}" for f, ln in chain[:3])))
        except ValueError:
            pass
    tt = sum(x[0] for x in rowsx)
    print(f"\ntop instructions by {sys.argv[6]} ({tt} samples)")
    for s_, off, txt, ch in sorted(rowsx, reverse=True)[:30]:
        print(f"{s_ / tt * 100:6.2f}%  {off:06x} {txt[:60]:60s} {ch}")

"""Launches per replayed epoch from an ncu launch list of tools/prof_graphed_epochs.py
(ncu --metrics gpu__time_duration.sum --clock-control none --csv): the span between the last two tc_kernel launches is
one config-4 epoch, the span covering the last six wan_scalars_kernel launches one config-5 minimax epoch.
Usage: python tools/epoch_launch_counts.py launches.csv"""
import csv
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
ks = [(r[ki], float(r[vi].replace(",", "")) / 1e3) for r in rows[1:]]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    name = name.replace("at::native::", "").replace("<unnamed>::", "")
    return name[:70]


def show(title, seg):
    print(f"{title}: {len(seg)} launches, {sum(t for _, t in seg):.1f} us of kernel time (ncu: cold caches, serialised)")
    for nm, t in seg:
        print(f"    {t:7.2f} us  {short(nm)}")


tc = [i for i, (nm, _) in enumerate(ks) if "tc_kernel" in nm]
first_wan = next((i for i, (nm, _) in enumerate(ks) if "wan_kernel" in nm), len(ks))
tc = [i for i in tc if i < first_wan]
if len(tc) >= 2:
    a, b = tc[-2], tc[-1]
    # an epoch starts at the launch after the previous epoch's Adam step; rotate so that it starts there
    adam = [i for i in range(a, b) if "adam" in ks[i][0]]
    s = adam[-1] + 1 if adam else a
    show("config 4 epoch (QHO_2D PINN + Adam)", ks[s:s + (b - a)])
sc = [i for i, (nm, _) in enumerate(ks) if "wan_scalars_kernel" in nm]
if len(sc) >= 13:
    a, b = sc[-13], sc[-7]          # one epoch earlier than the last, so that the span is whole
    seg = ks[a:b]
    show("config 5 minimax epoch (5 critic + 1 solution update), from a wan_scalars launch to the sixth after it", seg)

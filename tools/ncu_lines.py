#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`."""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur_file = None; hdr = None
inst = collections.Counter(); samp = collections.Counter(); src = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] in ("Function Name", "Kernel Name"): continue
    if hdr is None or r[0] == "": continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    i_inst = hdr.index("Instructions Executed"); i_s = hdr.index("# Samples")
    try:
        a = int(r[i_inst]); b = int(r[i_s])
    except ValueError:
        continue
    inst[(cur_file, ln)] += a; samp[(cur_file, ln)] += b; src[(cur_file, ln)] = r[1].strip()
ti = sum(inst.values()); ts = sum(samp.values())
print(f"total warp instructions {ti/1e9:.3f} G, samples {ts}")
print("== by instructions")
for k, v in inst.most_common(top):
    print(f"{k[0]}:{k[1]:<5d} inst {100*v/ti:5.1f}%  samp {100*samp[k]/max(ts,1):5.1f}%  {src[k][:110]}")
print("== by samples")
for k, v in samp.most_common(top // 2):
    print(f"{k[0]}:{k[1]:<5d} samp {100*v/max(ts,1):5.1f}%  inst {100*inst[k]/ti:5.1f}%  {src[k][:110]}")
# regions of search.cu
if len(sys.argv) > 3:
    import json
    regs = json.loads(sys.argv[3])
    for name, (lo, hi, f) in regs.items():
        a = sum(v for k, v in inst.items() if k[0] == f and lo <= k[1] <= hi); b = sum(v for k, v in samp.items() if k[0] == f and lo <= k[1] <= hi)
        print(f"{name:28s} inst {100*a/ti:5.1f}%  samp {100*b/max(ts,1):5.1f}%")

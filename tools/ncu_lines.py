"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line: share of warp-stall samples,
share of executed instructions, top stall reasons.   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | python tools/ncu_lines.py [N]"""
import csv, sys
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
hdr = None; fname = ""; data = []
for r in rows:
    if r and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0]: continue
    idx = {h: i for i, h in enumerate(hdr)}
    try:
        samp = int(r[idx["# Samples"]]); inst = int(r[idx["Instructions Executed"]])
    except Exception:
        continue
    st = {h: int(r[i] or 0) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h and r[i] not in ("", "-")}
    data.append((samp, inst, fname, r[0], r[1], st))
tot = sum(d[0] for d in data) or 1; toti = sum(d[1] for d in data) or 1
print(f"samples {tot}  warp instructions {toti}")
agg = {}
for d in data:
    for k, v in d[5].items(): agg[k] = agg.get(k, 0) + v
print("stall totals:", ", ".join(f"{k[6:]} {v/tot*100:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for samp, inst, f, ln, src, st in sorted(data, key=lambda d: -d[0])[:n]:
    top = ", ".join(f"{k[6:]} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"{samp/tot*100:5.1f}% smp {inst/toti*100:5.1f}% ins  {f}:{ln:>4s} {src.strip()[:90]:90s} | {top}")

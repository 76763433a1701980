"""Summarise `ncu --page source --csv` output: the most-sampled SASS instructions of
each kernel section with their dominant stall reasons.  usage: ncu_top.py file.csv [section] [topn]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        sections.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) >= len(cur["hdr"]):
        cur["data"].append(r)
print(len(sections), "sections")
sec = sections[want]
hdr, data = sec["hdr"], sec["data"]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp] or 0) for r in data)
print(sec["name"][:60], "instructions", len(data), "samples", tot)
agg = {}
for r in data:
    for c in stall_cols:
        if r[c]:
            agg[hdr[c][6:]] = agg.get(hdr[c][6:], 0) + int(r[c])
print("stall totals:", dict(sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
top = sorted(enumerate(data), key=lambda x: -int(x[1][isamp] or 0))[:topn]
for i, r in sorted(top):
    st = {hdr[c][6:]: int(r[c]) for c in stall_cols if r[c] and int(r[c]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{i:5d} {int(r[isamp]):6d} {r[iex]:>8s}  {r[isrc].strip()[:78]:78s} {st}")

"""developer tool: per-kernel hot spots of an exported `ncu --page source --csv --print-source sass` file"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 22
ks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1][:40], "hdr": None, "rows": []}; ks.append(cur)
    elif cur is not None and cur["hdr"] is None and r and r[0] == "Address": cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and r: cur["rows"].append(r)
for k in ks:
    h = k["hdr"]; ie = h.index("Instructions Executed"); ns = h.index("# Samples"); sm = h.index("L1 Wavefronts Shared")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[ie] or 0) for r in k["rows"]); ts = sum(int(r[ns] or 0) for r in k["rows"])
    print("=====", k["name"], "instr", tot, "samples", ts, "smem wf", sum(int(r[sm] or 0) for r in k["rows"]))
    agg = {}
    for r in k["rows"]:
        for i in stall_cols: agg[h[i]] = agg.get(h[i], 0) + int(r[i] or 0)
    print("  stalls:", sorted(((v, n) for n, v in agg.items()), reverse=True)[:8])
    for idx, r in sorted(enumerate(k["rows"]), key=lambda t: -int(t[1][ns] or 0))[:topn]:
        st = sorted([(int(r[i] or 0), h[i][6:]) for i in stall_cols], reverse=True)[:2]
        print(f"{idx:5d} {int(r[ns]):7d} {100*int(r[ns])/ts:5.2f}% ex{int(r[ie] or 0)/1e6:7.1f}M wf{int(r[sm] or 0)/1e6:7.1f}M {r[1][:58]:58s} {st}")

"""Summarise `ncu --page source --csv --print-source sass` output: per kernel, instruction mix by opcode
(share of executed warp instructions and of stall samples) and the hottest instructions."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 24
funcs, cur, hdr = [], None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; funcs.append(cur); hdr = None; continue
    if r and r[0] == "Address":
        hdr = r; cur["hdr"] = r; continue
    if cur is not None and hdr is not None and len(r) == len(hdr):
        cur["rows"].append(r)
for f in funcs:
    h = f["hdr"]; ie = h.index("Instructions Executed"); isrc = h.index("Source"); ism = h.index("Warp Stall Sampling (All Samples)")
    tot = sum(int(r[ie] or 0) for r in f["rows"]); ts = sum(int(r[ism] or 0) for r in f["rows"]) or 1
    print("=====", f["name"][:100]); print("sass instrs", len(f["rows"]), "executed warp instrs", tot, "stall samples", ts)
    ops, smp = collections.Counter(), collections.Counter()
    for r in f["rows"]:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[isrc]); op = m.group(2) if m else "?"
        ops[op] += int(r[ie] or 0); smp[op] += int(r[ism] or 0)
    for op, c in ops.most_common(top):
        print(f"  {op:10s} {c / tot * 100:5.1f}% instr  {smp[op] / ts * 100:5.1f}% samples")
    print("  -- hottest instructions by stall samples")
    for r in sorted(f["rows"], key=lambda r: -int(r[ism] or 0))[:12]:
        print(f"     {int(r[ism]) / ts * 100:4.1f}%  exec {r[ie]:>9s}  {r[isrc].strip()[:90]}")

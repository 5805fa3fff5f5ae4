"""Per-source-line executed-instruction totals from `ncu --page source --print-source cuda,sass --csv`.
    python tools/ncu_lines.py report.ncu-rep [Mpx per launch] [top N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; mpx = float(sys.argv[2]) if len(sys.argv) > 2 else 9.8304; top = int(sys.argv[3]) if len(sys.argv) > 3 else 70
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fname, hdr, rows = None, None, []
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0] not in ("", "Function Name") and len(r) == len(hdr):
        ie = hdr.index("Instructions Executed"); ism = hdr.index("Warp Stall Sampling (All Samples)")
        try: rows.append((fname, int(r[0]), r[1].strip(), int(r[ie] or 0), int(r[ism] or 0)))
        except ValueError: pass
tot = sum(x[3] for x in rows); ts = sum(x[4] for x in rows) or 1
print(f"total thread-instr/px {tot * 32 / mpx / 1e6:.1f}")
for f, ln, txt, e, s in sorted(rows, key=lambda x: -x[3])[:top]:
    print(f"{e * 32 / mpx / 1e6:7.1f}  {s / ts * 100:5.1f}%  {f}:{ln}  {txt[:110]}")

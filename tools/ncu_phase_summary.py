"""Key raw metrics + per-barrier-segment instruction breakdown from an .ncu-rep (run where ncu is on PATH).
    python tools/ncu_phase_summary.py report.ncu-rep [megapixels per launch]"""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
mpx = float(sys.argv[2]) if len(sys.argv) > 2 else 9.8304
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__t_sectors_srcunit_tex_op_red.sum", "sm__cycles_elapsed.max"]
for v in rows[2:]:
    print("=====", v[h.index("Kernel Name")][:110])
    for k in keys:
        if k in h:
            print(f"  {k:72s} {v[h.index(k)]:>18s} {u[h.index(k)]}")
    for i, k in enumerate(h):
        if "average_warps_issue_stalled" in k and "not_issued" not in k and float(v[i] or 0) > 0.05:
            print(f"  stall {k.split('stalled_')[1].split('_per_')[0]:28s} {float(v[i]):.3f} warps per issue")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
hdr, data = None, []
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        data.append(r)
if hdr:
    ie, isrc, ism = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
    tot = sum(int(r[ie] or 0) for r in data); ts = sum(int(r[ism] or 0) for r in data) or 1
    segs, cur = [], []
    for r in data:
        cur.append(r)
        if "BAR." in r[isrc]:
            segs.append(cur); cur = []
    segs.append(cur)
    print(f"  total thread-instr/px {tot * 32 / (mpx * 1e6):.1f}")
    for i, s in enumerate(segs):
        e = sum(int(r[ie] or 0) for r in s); smp = sum(int(r[ism] or 0) for r in s)
        if e * 50 < tot:
            continue
        ops = collections.Counter()
        for r in s:
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[isrc]); ops[m.group(2) if m else "?"] += int(r[ie] or 0)
        print(f"  segment {i}: {len(s)} sass, {e / tot * 100:.1f}% of instr, {smp / ts * 100:.1f}% of samples, {e * 32 / (mpx * 1e6):.1f} thread-instr/px")
        print("     " + " ".join(f"{k}:{c * 32 / (mpx * 1e6):.1f}" for k, c in ops.most_common(24)))
    print("  hottest by samples:")
    for r in sorted(data, key=lambda r: -int(r[ism] or 0))[:14]:
        print(f"     {int(r[ism]) / ts * 100:4.1f}%  exec {r[ie]:>9s}  {r[isrc].strip()[:100]}")

"""tools/prof_lines.py REPORT.ncu-rep [N] -- top CUDA source lines by warp-stall samples, with stall reasons."""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if "# Samples" in r][0]
hdr = rows[hi]
stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = 0; lines = []
for r in rows[hi + 1:]:
    if len(r) <= iex or not r[0].strip().isdigit(): continue
    try: n = int(r[isamp]); ex = int(r[iex])
    except ValueError: continue
    top = sorted([(int(r[i] or 0), h[6:]) for i, h in stalls], reverse=True)[:3]
    lines.append((n, int(r[0]), ex, r[1].strip()[:88], [t for t in top if t[0]])); tot += n
print("total samples", tot, " total instructions", sum(l[2] for l in lines))
for n, ln, ex, src, top in sorted(lines, reverse=True)[:topn]:
    print("%5d %5.1f%% L%-4d ex=%-8d %-88s %s" % (n, 100.0 * n / max(tot, 1), ln, ex, src, top))

"""tools/sass_opcodes.py -- SASS opcode counts per kernel family of aligner_b200/libaligner_b200.so (writes profiles/r02_sass_opcodes.txt).
No GPU needed (cuobjdump reads the embedded sm_100a cubins)."""
import re, subprocess, sys
from collections import Counter, OrderedDict
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
lib = ROOT / "aligner_b200" / "libaligner_b200.so"
out = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
FAM = OrderedDict([("mas_kernel", r"mas_kernel"), ("neg_cent FFMA kernels", r"albnc\d+(ota|gaussian)_kernel"), ("nc_tc_kernel (first-generation tcgen05)", r"nc_tc_kernel"),
                   ("nc_prep_kernel", r"nc_prep_kernel"), ("nc_v2_kernel (second-generation tcgen05)", r"nc_v2_kernel")])
MARK = re.compile(r"^(UTC|LDTM|STTM|UTMA|UBLKCP|STAS|SYNCS|LDGSTS|ACQBULK|FENCE|MEMBAR|BAR|ELECT|ERRBAR|CGAERRBAR|MUFU|REDUX|UCGABAR|R2UR|NANOSLEEP|BPT|CCTL)")
counts = {k: Counter() for k in FAM}
other = Counter()
cur = None
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = None
        for k, pat in FAM.items():
            if re.search(pat, m.group(1)): cur = k; break
        if cur is None: other[m.group(1)] += 1
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", ln)
    if m and cur: counts[cur][m.group(1)] += 1
lines = ["# SASS opcode counts of aligner_b200/libaligner_b200.so (cuobjdump -sass, sm_100a; tools/sass_opcodes.py), all instances of a kernel family summed.",
         "# Blackwell-native markers: UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA load; .2D dense tiles, .3D the pre-skewed boxes of the",
         "# 4-frame-lag form and the score kernel's operand boxes), UBLKCP = cp.async.bulk, STAS = st.async (cluster), SYNCS = mbarrier, UTCBAR = tcgen05.commit,",
         "# LDGSTS = cp.async, BPT.TRAP = the bounded waits' give-up path.", ""]
for k, c in counts.items():
    tot = sum(c.values())
    if not tot: continue
    lines.append("== %s  (%d instructions)" % (k, tot))
    lines.append("   markers: " + "  ".join("%s x%d" % (o, n) for o, n in sorted(c.items()) if MARK.match(o)))
    lines.append("   top 40:  " + "  ".join("%s x%d" % (o, n) for o, n in c.most_common(40)))
    lines.append("")
if other: lines.append("# other functions: " + ", ".join(sorted(other)[:8]))
(ROOT / "profiles" / "r02_sass_opcodes.txt").write_text("\n".join(lines) + "\n")
print("\n".join(l[:200] for l in lines[:14]))

"""Instruction histogram of the headline kernel from the in-tree library (cuobjdump -sass): the evidence that
fir_tc_kernel is tcgen05 / TMEM / bulk-TMA code (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP =
cp.async.bulk, UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc), per /opt/skills/guides/B200_PROFILING.md.

    python scripts/sass_histogram.py > profiles/sass_fir_tc.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "futuresdr_b200", "libb200sdr.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
want = sys.argv[1] if len(sys.argv) > 1 else "fir_tc_kernel"
cur, hist = None, {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1) if want in m.group(1) else None
        if cur:
            hist[cur] = collections.Counter()
        continue
    if cur:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m:
            hist[cur][m.group(1)] += 1
            if m.group(1) in ("UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR", "UTCATOMSWS", "SYNCS", "UTMALDG", "ELECT"):
                hist[cur][m.group(1) + m.group(2)] += 1
print(f"# cuobjdump -sass {os.path.relpath(so, ROOT)} | functions matching '{want}' (sm_100a)")
KEY = ("UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR", "UTCATOMSWS", "SYNCS", "ELECT", "UTMALDG")
for fn, h in hist.items():
    print(f"\n## {fn}\ntotal instructions: {sum(v for k, v in h.items() if '.' not in k)}")
    print("Blackwell-specific:")
    for k in sorted(h):
        if k.split(".")[0] in KEY:
            print(f"  {k:40s} {h[k]}")
    print("top 25 mnemonics:")
    for k, v in collections.Counter({k: v for k, v in h.items() if "." not in k}).most_common(25):
        print(f"  {k:16s} {v}")

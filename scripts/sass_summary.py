#!/usr/bin/env python
"""Counts of the SASS mnemonics that prove what each kernel of libcmh_b200.so runs on (cuobjdump -sass; no GPU needed):
UTCIMMA / UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (bulk-copy / TMA
engine), SYNCS = mbarrier, POPC / MATCH / VOTE / ATOMS = the integer paths.   python scripts/sass_summary.py > profiles/sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "clip-based-cross-modal-hashing_b200", "libcmh_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ("UTCIMMA", "UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "POPC", "MATCH", "VOTE", "ATOMS", "RED", "MUFU", "SHFL", "HMMA", "IMMA")
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", name)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        per[cur]["_total"] += 1
        for k in KEYS:
            if op.startswith(k):
                per[cur][k] += 1
print(f"# cuobjdump -sass {os.path.relpath(so, ROOT)} (sm_100a); instruction counts per kernel; only non-zero columns shown")
for name, c in per.items():
    cols = " ".join(f"{k}={c[k]}" for k in KEYS if c[k])
    print(f"{name[:100]:100s} total={c['_total']:6d} {cols}")

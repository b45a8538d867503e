"""SASS evidence for profiles/: per kernel of libctp.so the counts of the Blackwell-native mnemonics and the first occurrence of each.
    python tests/sass_excerpts.py [kernel-name regex] > profiles/rN_sass_excerpts.txt      (runs without a GPU: cuobjdump only)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
LIB = os.path.join(ROOT, "chatttsplus_b200", "_C", "libctp.so")
WANT = ["UTCHMMA", "UTCBAR", "UTCATOM", "LDTM", "STTM", "UTMALDG", "UTMACCTL", "UBLKCP", "UBLKPF", "UTMAPF", "ACQBULK", "SYNCS", "ELECT", "REDG", "HMMA",
        "UCGABAR", "CCTL", "MAPA", "ATOMS"]
pat = re.compile(sys.argv[1]) if len(sys.argv) > 1 else re.compile("gemm_tcgen05|k_mlp_fused|k_attn_decode_tma|k_attn_prefill_mma|k_sample|k_istft|k_dwconv")
out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
print("# SASS evidence (cuobjdump -sass chatttsplus_b200/_C/libctp.so, sm_100a, built from this tree by `python -m chatttsplus_b200.build`)")
print("# For every hot kernel: counts of the Blackwell-native mnemonics (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM,")
print("# TMA -> UTMALDG / UBLKCP, cp.async.bulk.prefetch.L2 -> UBLKPF, tcgen05.commit -> UTCBAR, barrier.cluster -> UCGABAR, legacy mma.sync -> HMMA)")
print("# and the first occurrence of each.\n")
name, lines = None, []


def flush():
    if not name or not pat.search(name):
        return
    cnt = collections.Counter()
    first = {}
    for ln in lines:
        m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if not m:
            continue
        op = m.group(1)
        for w in WANT:
            if op.startswith(w):
                cnt[w] += 1
                first.setdefault(w, ln.strip().split("/*", 2)[0] if False else ln.rstrip())
    print(f"\n== {name}")
    print("   counts: " + ", ".join(f"{k} {v}" for k, v in sorted(cnt.items())))
    for w, ln in first.items():
        print(f"   first {w}:")
        print("      " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", ln.strip()))


for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        flush()
        name, lines = m.group(1), []
    else:
        lines.append(ln)
flush()

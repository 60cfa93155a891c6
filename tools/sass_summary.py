"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md): UTCHMMA (tcgen05.mma),
LDTM / STTM (tcgen05.ld / .st), UTMALDG / UTMASTG (TMA tensor loads / stores), UTCBAR (tcgen05.commit), HMMA (mma.sync).
usage: python tools/sass_summary.py > profiles/sass_summary.txt   (runs cuobjdump -sass on the in-tree libopus_b200.so)"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "opus_pllm_b200", "libopus_b200.so")
OPS = ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "SYNCS", "HMMA", "LDGSTS")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
counts, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::|opus::", "", cur).split("(")[0]
        cur = re.sub(r"^void ", "", cur)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_n"] += 1
        for k in OPS:
            if op.startswith(k):
                counts[cur][k] += 1
print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a), instruction counts per kernel")
print(f"{'kernel':78s} {'instr':>7s} " + " ".join(f"{k:>8s}" for k in OPS))
tot = collections.Counter()
for name, c in counts.items():
    print(f"{name[:78]:78s} {c['_n']:7d} " + " ".join(f"{c[k]:8d}" for k in OPS))
    tot.update(c)
print(f"{'TOTAL':78s} {tot['_n']:7d} " + " ".join(f"{tot[k]:8d}" for k in OPS))

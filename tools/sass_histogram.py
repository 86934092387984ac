"""Opcode evidence per kernel of libdbn_b200.so (cuobjdump -sass): the Blackwell-native instructions each kernel
contains -- UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTMALDG/UTMASTG (TMA tensor copies), UBLKCP (bulk
copies, here DSMEM), SYNCS (mbarrier), UTCBAR (tcgen05.commit), REDG/ATOMG (the fixed-point statistics), plus the
absence of HMMA (legacy mma.sync).      python tools/sass_histogram.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
lib = os.path.join(ROOT, "deep-belief-network_b200", "libdbn_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTC[A-Z]*MMA(?:\.2CTA)?|UTCBAR(?:\.2CTA)?(?:\.MULTICAST)?|LDTM(?:\.[0-9x]+)?|STTM|UTMALDG(?:\.[0-9]D)?(?:\.2CTA)?|UTMASTG|UBLKCP(?:\.S\.S)?|UTMAPF|"
                 r"SYNCS(?:\.[A-Z_0-9]+)*|HMMA|REDG?(?:\.E)?(?:\.ADD)?(?:\.64)?|ATOMG?|MEMBAR(?:\.[A-Z.]+)?|UCGABAR_[A-Z]+|MUFU\.(?:EX2|RCP))")
fn = None
hist = collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"\(.*", "", fn)
        hist[fn] = collections.Counter()
        continue
    if fn is None:
        continue
    m = pat.search(line.split("/*")[1] if "/*" in line and line.strip().startswith("/*") else line)
    if m and ";" in line:
        op = m.group(1)
        op = re.sub(r"^(SYNCS)\..*", r"\1", op)
        hist[fn][op] += 1
print("# cuobjdump -sass deep-belief-network_b200/libdbn_b200.so : Blackwell-native opcodes per kernel (count of static instructions)")
for fn, h in hist.items():
    if not h:
        continue
    print("\n%s" % fn)
    print("   " + "  ".join("%s x%d" % (k, v) for k, v in sorted(h.items())))
legacy = sum(h.get("HMMA", 0) for h in hist.values())
print("\nlegacy tensor path (HMMA = mma.sync / wmma) instructions in the library: %d" % legacy)

"""cuobjdump -sass libsib200.so | python scripts/sass_histogram.py > profiles/rNN_sass_mnemonic_histogram.txt
Per-kernel counts of the tcgen05 (UTC*), TMEM (LDTM), TMA (UTMA*), mbarrier (SYNCS) and setmaxnreg
mnemonics; any legacy HMMA would show up too."""
import re
import subprocess
import sys
from collections import Counter, OrderedDict

PAT = re.compile(r"\b(UTC[A-Z0-9]+(?:\.[A-Z0-9_]+)*|LDTM(?:\.[A-Za-z0-9_]+)*|UTMA[A-Z]+(?:\.[A-Z0-9_]+)*|"
                 r"SYNCS\.[A-Z0-9_.]+|USETMAXREG\.[A-Z.]+|HMMA\S*)")
cur, hist = None, OrderedDict()
for line in sys.stdin:
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = Counter()
        continue
    if cur is not None:
        for mn in PAT.findall(line):
            hist[cur][mn] += 1
names = subprocess.run(["c++filt"] + list(hist.keys()), capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump -sass sota_imagenet_b200/libsib200.so | python scripts/sass_histogram.py")
print("# tcgen05 (UTC*), TMEM (LDTM), TMA (UTMA*), mbarrier (SYNCS, summed), setmaxnreg per kernel; HMMA would be the legacy path")
tot = Counter()
for (k, c), n in zip(hist.items(), names):
    if not c:
        continue
    tot.update(c)
    short = re.sub(r"\(.*", "", n).replace("void sib::", "")
    sy = sum(b for a, b in c.items() if a.startswith("SYNCS"))
    print("%-44s %s  SYNCS x%d" % (short[:44], "  ".join("%s x%d" % (a, b) for a, b in sorted(c.items())
                                                          if not a.startswith("SYNCS")), sy))
print("# total: " + "  ".join("%s x%d" % (a, b) for a, b in sorted(tot.items()) if not a.startswith("SYNCS")))
print("# HMMA (legacy mma.sync) instructions: %d" % sum(b for a, b in tot.items() if a.startswith("HMMA")))

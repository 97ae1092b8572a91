"""Counts the Blackwell tensor-path SASS mnemonics per kernel of librealise_b200.so (cuobjdump -sass):
UTCHMMA(.2CTA) = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UTMAREDG = TMA load/store/reduce (".MULTICAST"
= multicast through the cluster), UTCBAR = tcgen05.commit, UTCATOMSWS = TMEM alloc.   python tools/sass_summary.py > profiles/rNN_sass_tensor_path.txt"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "realise_b200/librealise_b200.so"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
fn, counts = None, collections.OrderedDict()
pat = re.compile(r"\b(UTC[A-Z]*MMA[.\w]*|LDTM[.\w]*|STTM[.\w]*|UTMALDG[.\w]*|UTMASTG[.\w]*|UTMAREDG[.\w]*|UTCBAR[.\w]*|UTCATOMSWS[.\w]*|UBLKCP[.\w]*|HMMA[.\w]*)")
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        counts[fn] = collections.Counter()
        continue
    if fn:
        m = pat.search(line)
        if m:
            counts[fn][m.group(1)] += 1
for fn, c in counts.items():
    if c:
        name = re.sub(r"\(anonymous namespace\)::", "", demangle(fn))
        name = re.sub(r"\(CUtensorMap_st.*", "", name).replace("void ", "")
        print(f"{name}\n    " + ", ".join(f"{k} x{v}" for k, v in sorted(c.items())))

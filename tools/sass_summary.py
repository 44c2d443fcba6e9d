"""SASS evidence of the Blackwell-native path: per kernel of libmrgan.so, the number of tcgen05 / TMEM / TMA instructions
(cuobjdump -sass; the PTX names never appear in SASS: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM,
cp.async.bulk.tensor -> UTMALDG/UTMASTG).   python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "mr_gan_b200", "libmrgan.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
MN = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "MUFU", "LDG", "STG"]
per, name, n = collections.OrderedDict(), None, collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        per[name] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1)
        per[name]["total"] += 1
        for k in MN:
            if op.startswith(k):
                per[name][k] += 1
print("libmrgan.so (sm_100a) -- instructions per kernel; tcgen05.mma = UTC*MMA, tcgen05.ld/st = LDTM/STTM, TMA = UTMALDG/UTMASTG")
print("%-72s %6s " % ("kernel", "instr") + " ".join("%7s" % k for k in MN))
tot = collections.Counter()
for k, c in per.items():
    print("%-72s %6d " % (k[:72], c["total"]) + " ".join("%7d" % c[m] for m in MN))
    tot.update(c)
print("%-72s %6d " % ("TOTAL", tot["total"]) + " ".join("%7d" % tot[m] for m in MN))
print("legacy tensor path (HMMA = mma.sync / wmma): %d instructions" % tot["HMMA"])

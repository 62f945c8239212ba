"""Blackwell-native evidence from the shipped library (run on the CPU box):
python tools/sass_histogram.py > profiles/r02_sass_histogram.txt
Counts the SASS mnemonics that prove tcgen05 / TMEM / TMA use, per kernel of cobevt_b200/libcobevt_sm100.so."""
import collections, os, re, subprocess, sys
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cobevt_b200", "libcobevt_sm100.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMACMDFLUSH", "SYNCS", "MUFU.EX2", "MUFU.RCP", "FFMA2", "FADD2", "FMUL2",
        "LDGSTS", "HMMA", "ELECT", "USETMAXREG"]
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*", "", name)
        cur = per.setdefault(name, collections.Counter())
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["_total"] += 1
        for k in KEYS:
            if op.startswith(k):
                cur[k] += 1
print(f"# SASS mnemonic counts per kernel of {os.path.basename(lib)} (cuobjdump -sass; sm_100a)")
print("# UTCHMMA = tcgen05.mma kind::f16, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = cp.async.bulk.tensor load/store,")
print("# UBLKCP = cp.async.bulk, SYNCS = mbarrier ops, LDGSTS = cp.async, HMMA = legacy mma.sync (must be 0)")
tot = collections.Counter()
for name, c in per.items():
    if c["_total"] < 50:
        continue
    cols = " ".join(f"{k}={c[k]}" for k in KEYS if c[k])
    print(f"{name[:70]:70s} instr={c['_total']:6d}  {cols}")
    tot.update(c)
print("TOTAL " + " ".join(f"{k}={tot[k]}" for k in KEYS))

"""Per-kernel counts of the Blackwell tensor-core / TMA / TMEM instructions in libgloria_b200.so (cuobjdump -sass).
usage: python scripts/sass_counts.py [lib] > profiles/r02_sass_counts.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "gloria_nlp_project_b200", "libgloria_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
names = dict(zip(re.findall(r"Function : (\S+)", out), demangle))
PAT = [("UTCHMMA", r"\bUTCHMMA"), ("UTCQMMA/UTCOMMA", r"\bUTC[QO]MMA"), ("UTCHMMA.2CTA", r"UTCHMMA\.2CTA"), ("LDTM", r"\bLDTM"),
       ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("UTMAPF", r"\bUTMAPF|UTMACCTL"),
       ("UTCBAR", r"\bUTCBAR"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("UCGABAR (cluster barrier)", r"\bUCGABAR"),
       ("HMMA/IMMA (legacy mma.sync)", r"\b[HI]MMA\b")]
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = names.get(m.group(1), m.group(1))
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for tag, pat in PAT:
        if re.search(pat, line):
            counts[cur][tag] += 1
print(f"# {os.path.basename(lib)}: sm_100a SASS instruction counts per kernel (cuobjdump -sass), kernels with tensor-core /"
      " TMA / TMEM instructions only")
print("# build id:", subprocess.run([sys.executable, "-c", "import ctypes,sys;l=ctypes.CDLL(sys.argv[1]);l.gloria_b200_build_id.restype=ctypes.c_char_p;print(l.gloria_b200_build_id().decode())", lib], capture_output=True, text=True).stdout.strip())
tot = collections.Counter()
for k, c in counts.items():
    if not any(c[t] for t in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG")):
        continue
    short = k.split("(CUtensorMap")[0].split("(const ")[0].replace("gloria::tc::", "").replace("(bool)", "")
    print(f"{short[:90]:90s} " + "  ".join(f"{t}={c[t]}" for t, _ in PAT if c[t]))
    tot.update(c)
print("TOTAL".ljust(90), "  ".join(f"{t}={tot[t]}" for t, _ in PAT if tot[t]))

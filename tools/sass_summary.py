"""Mnemonic counts per tensor-core kernel of libpcrl_b200.so (runs without a GPU):
   python tools/sass_summary.py > profiles/rNN_sass_tensor_kernels.md
UTCHMMA = tcgen05.mma, UTMALDG = cp.async.bulk.tensor (TMA), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,
HMMA without the UTC prefix = legacy mma.sync (must be 0)."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "pcrlv2_b200", "libpcrl_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
print("# SASS evidence of the tensor-core kernels in pcrlv2_b200/libpcrl_b200.so\n")
print("`cuobjdump -sass pcrlv2_b200/libpcrl_b200.so`, mnemonic counts per kernel (tools/sass_summary.py).\n")
print("| kernel | UTCHMMA | UTMALDG | LDTM | UTCBAR | legacy HMMA | STG .256 |\n|---|---|---|---|---|---|---|")
tot = collections.Counter()
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    dem = re.sub(r"\(.*", "", subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip())
    c = dict(UTCHMMA=len(re.findall(r"\bUTCHMMA", f)), UTMALDG=len(re.findall(r"\bUTMALDG", f)),
             LDTM=len(re.findall(r"\bLDTM", f)), UTCBAR=len(re.findall(r"\bUTCBAR", f)),
             HMMA=len(re.findall(r"(?<![A-Z])HMMA", f)), STG256=len(re.findall(r"STG\.E[^ ]*\.256", f)))
    if c["UTCHMMA"] or c["UTMALDG"] or c["LDTM"]:
        tot.update(c)
        print(f"| `{dem}` | {c['UTCHMMA']} | {c['UTMALDG']} | {c['LDTM']} | {c['UTCBAR']} | {c['HMMA']} | {c['STG256']} |")
print(f"| **total** | {tot['UTCHMMA']} | {tot['UTMALDG']} | {tot['LDTM']} | {tot['UTCBAR']} | {tot['HMMA']} | {tot['STG256']} |")
print(f"\nWhole library: {len(funcs)} kernels.")

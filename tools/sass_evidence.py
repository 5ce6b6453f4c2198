"""Writes profiles/<name>: SASS mnemonic counts and excerpts of the product library (cuobjdump -sass), the evidence that the kernels use the
instructions DESIGN.md names -- DMMA (mma.sync.m8n8k4.f64), UBLKCP (cp.async.bulk, TMA), SYNCS (mbarrier), LDGSTS (cp.async), system-scope
stores / fences on peer memory.     usage: python tools/sass_evidence.py [profiles/r2_sass_evidence.md]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "spectra_b200", "lib", "libspectra_b200.so")
dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_evidence.md")
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
WANT = r"(compress_dmma_kernelILi3E|sell_step_dot_kernelILb1ELb1ELi\dE|peer_allreduce_kernel|panel_kernelILi4ELi2ELb0E|peer_push_kernel|sell_plain_kernelILi512ELb1E|sell_step_kernelILi512ELb1ELb1E)"
PATS = [("DMMA", r"\bDMMA"), ("UBLKCP", r"UBLKCP"), ("SYNCS", r"SYNCS"), ("LDGSTS", r"LDGSTS"), ("sys-scope LD/ST", r"(LDG|STG|LD|ST)\.[\w.]*SYS"), ("MEMBAR.SC.SYS", r"MEMBAR\.SC\.SYS")]
out = ["# SASS evidence (cuobjdump -sass spectra_b200/lib/libspectra_b200.so, sm_100a)\n",
       "PTX names do not appear in SASS: `mma.sync.m8n8k4.f64` -> `DMMA`, `cp.async.bulk` -> `UBLKCP`, `mbarrier.*` -> `SYNCS`, `cp.async` -> `LDGSTS`,",
       "`st.release.sys` / `ld.acquire.sys` on peer memory -> `STG...STRONG.SYS` / `LDG...STRONG.SYS`, `fence.sys` -> `MEMBAR.SC.SYS`.\n",
       "| kernel | " + " | ".join(n for n, _ in PATS) + " |", "|---|" + "---:|" * len(PATS)]
ex = []
for f in funcs[1:]:
    name = f.split("\n", 1)[0]
    m = re.search(WANT, name)
    if not m:
        continue
    short = m.group(1)
    out.append(f"| `{short}` | " + " | ".join(str(len(re.findall(p, f))) for _, p in PATS) + " |")
    for _, p in PATS:
        ls = [l.strip() for l in f.split("\n") if re.search(p, l)]
        if ls:
            ex.append(f"{short}: {ls[0][:140]}")
out += ["\nExcerpts (first occurrence per kernel and mnemonic):\n```"] + ex + ["```"]
open(dst, "w").write("\n".join(out) + "\n")
print("wrote", dst)

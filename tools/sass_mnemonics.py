"""SASS mnemonic counts per kernel of the in-tree library (`cuobjdump -sass`, no GPU needed):
    python tools/sass_mnemonics.py [lib.so] [commit] > profiles/r02_sass_mnemonics.txt
tcgen05.mma = UTCHMMA, tcgen05.ld = LDTM, TMA load / store = UTMALDG / UTMASTG, tcgen05.commit = UTCBAR."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "multimodal-vae_b200", "libmvae_b200.so")
commit = sys.argv[2] if len(sys.argv) > 2 else subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"],
                                                             capture_output=True, text=True).stdout.strip()
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names, cur, body = [], None, {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        names.append(cur)
        body[cur] = []
        continue
    if cur:
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m:
            body[cur].append(m.group(1).strip())
demangled = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()


def short(d):
    d = re.sub(r"\(anonymous namespace\)::|mvae::|void ", "", d)
    d = re.sub(r"\(.*$", "", d)            # argument list
    return d.replace("(int)", "").replace("(bool)", "")


keep = re.compile(r"chain_kernel|dp_reduce_adam|gemm_kernel|tail_|poe_|textenc|textdec")
print("# SASS mnemonic counts per kernel of libmvae_b200.so (cuobjdump -sass, sm_100a, built from the tree at commit %s;" % commit)
print("# tools/sass_mnemonics.py).  tcgen05.mma = UTCHMMA, tcgen05.ld = LDTM, TMA load / store = UTMALDG / UTMASTG, tcgen05.commit =")
print("# UTCBAR, '2CTA' = instructions carrying the .2CTA modifier (cta_group::2 MMA / TMA / commit / TMEM allocation), F*2 = packed")
print("# fp32 pairs (FADD2 + FMUL2 + FFMA2), UCGABAR = cluster barrier.  chain_kernel<K0, K1, K2, pair>: kinds 0 FWD_BN, 1 FWD_STORE,")
print("# 2 BCE, 3 DGRAD_BN, 4 DGRAD_STORE.  ACQBULK / PREEXIT = griddepcontrol.wait / .launch_dependents (programmatic launches).")
for mangled, d in zip(names, demangled):
    if not keep.search(d):
        continue
    ins = body[mangled]
    op = [i.split()[1] if i.startswith("@") and len(i.split()) > 1 else i.split()[0] for i in ins if i]

    def n(pat):
        return sum(1 for o in op if re.match(pat, o))
    print("%-44s instr %6d  UTCHMMA %3d  LDTM %3d  UTMALDG %3d  UTMASTG %3d  UTCBAR %3d  2CTA %3d  MUFU %4d  F*2 %4d  UCGABAR %2d  ACQBULK %d  PREEXIT %d" % (
        short(d), len(ins), n(r"UTC[A-Z]*MMA"), n(r"LDTM"), n(r"UTMALDG"), n(r"UTMASTG"), n(r"UTCBAR"),
        sum(1 for o in op if ".2CTA" in o), n(r"MUFU"), n(r"(FADD2|FMUL2|FFMA2)"), n(r"UCGABAR"), n(r"ACQBULK"), n(r"PREEXIT")))

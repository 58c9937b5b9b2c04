"""Split `cuobjdump -sass lib.so` into per-kernel instruction streams (addresses / encodings stripped) for before/after diffs:
    python tools/sass_split.py lib.so outdir [name-filter]"""
import os, re, subprocess, sys
lib, out = sys.argv[1], sys.argv[2]
flt = sys.argv[3] if len(sys.argv) > 3 else "gemm_kernel"
os.makedirs(out, exist_ok=True)
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, buf = None, {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1) if flt in m.group(1) else None
        if cur:
            buf[cur] = []
        continue
    if cur:
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m:
            buf[cur].append(m.group(1).strip())
for name, ins in buf.items():
    key = re.search(r"gemm_kernelILi(\d)ELi(\d)EL[bi](\d)", name)
    fn = "gemm_%s_%s_%s.sass" % key.groups() if key else re.sub(r"\W", "_", name)[:80] + ".sass"
    open(os.path.join(out, fn), "w").write("\n".join(ins) + "\n")
    print(fn, len(ins))

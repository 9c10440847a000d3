"""SASS opcode histogram of libda_b200.so: tcgen05 / TMEM / TMA / bulk-copy / mbarrier opcodes per kernel (cuobjdump -sass).
    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections, os, re, subprocess, sys
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                  "unsupervised_domain_adaptation_object_detection_implementation_b200", "libda_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
pat = re.compile(r"\b(UTC[A-Za-z0-9_.]+|LDTM[A-Za-z0-9_.]*|STTM[A-Za-z0-9_.]*|UTMA[A-Za-z0-9_.]+|UBLKCP[A-Za-z0-9_.]*|UTMAPF[A-Za-z0-9_.]*|SYNCS[A-Za-z0-9_.]*)")
per, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur and "/*" in line:
        for op in pat.findall(line.split("/*")[1] if line.strip().startswith("/*") else line):
            per[cur][op] += 1
names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
total = collections.Counter()
for c in per.values():
    total.update(c)
print("# SASS opcode histogram of libda_b200.so (cuobjdump -sass, sm_100a): tcgen05 / TMEM / TMA / bulk-copy / mbarrier opcodes per kernel")
print("# total:", dict(sorted(total.items())))
for (mangled, c), name in zip(per.items(), names):
    if c:
        print(f"{name[:110]} :: " + ", ".join(f"{k} x{v}" for k, v in sorted(c.items())))

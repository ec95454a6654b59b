"""Opcode evidence: count the SASS mnemonics that prove tcgen05 / TMEM / TMA / mbarrier / PDL use, per kernel of
rmnet_b200/librmnet_b200.so (cuobjdump -sass).  usage: python tools/sass_table.py > profiles/rNN_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "rmnet_b200", "librmnet_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
OPS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "SYNCS", "ACQBULK", "REDUX", "MUFU.EX2",
       "HMMA", "FFMA", "LDG", "STG", "ATOMG", "RED", "SHFL", "BAR.SYNC", "ELECT", "F2FP"]
per = collections.OrderedDict()
name = None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("rmnet::(anonymous namespace)::", "").replace("void ", "")
        per[name] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1)
        per[name]["_total"] += 1
        for o in OPS:
            if op == o or op.startswith(o + ".") or (o == "MUFU.EX2" and op.startswith("MUFU.EX2")):
                per[name][o] += 1
arch = re.findall(r"arch = (sm_\w+)", txt)
print(f"# cuobjdump -sass rmnet_b200/librmnet_b200.so  (arch: {sorted(set(arch))}); counts of instructions per kernel")
for k, c in per.items():
    hits = " ".join(f"{o}={c[o]}" for o in OPS if c[o])
    print(f"{k}: total={c['_total']} {hits}")

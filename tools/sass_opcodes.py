"""SASS opcode summary of the in-tree library (cuobjdump -sass): per kernel the instruction count and the opcodes that prove the
Blackwell data-movement paths (UTMALDG = 2-D TMA tensor copy, UBLKCP = 1-D bulk copy, LDGSTS = cp.async, SYNCS = mbarrier
ops) and the packed FP32 arithmetic (FFMA2 / FADD2); no tensor-core opcodes by design.   usage: sass_opcodes.py > profiles/sass_opcodes_r02.txt"""
import collections, os, re, subprocess
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "opticalimageprocessor_b200", "liboip_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, per = None, collections.OrderedDict()
archs = set(re.findall(r"arch = (sm_\w+)", txt))
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        per[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        per[kern][m.group(1)] += 1
watch = ["UTMALDG", "UBLKCP", "LDGSTS", "SYNCS", "FFMA2", "FADD2", "DFMA", "DMUL", "DADD", "SHFL", "LOP3", "UTCMMA", "HMMA", "LDTM"]
print(f"{lib}: architectures {sorted(archs)}, {len(per)} kernels")
print(f"{'kernel':64s} {'instr':>8s} " + " ".join(f"{w:>8s}" for w in watch))
tot = collections.Counter()
for k, c in per.items():
    row = {w: sum(v for op, v in c.items() if op.split(".")[0] == w or op.startswith(w)) for w in watch}
    tot.update(row)
    print(f"{k[-64:]:64s} {sum(c.values()):8d} " + " ".join(f"{row[w]:8d}" for w in watch))
print(f"{'TOTAL':64s} {sum(sum(c.values()) for c in per.values()):8d} " + " ".join(f"{tot[w]:8d}" for w in watch))

"""decode the scheduling control bits of a kernel's SASS (cuobjdump -sass): per basic block the instruction count and the
sum of the encoded stall counts = cycles ONE warp needs for the block if no scoreboard wait ever blocks it.
usage: sass_stalls.py lib.so kernel_substring [min_block_instrs]"""
import re, subprocess, sys, collections
lib, key = sys.argv[1], sys.argv[2]
min_n = int(sys.argv[3]) if len(sys.argv) > 3 else 300
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
on, ins = False, []
for i, l in enumerate(txt):
    if "Function :" in l:
        on = key in l
        continue
    if not on: continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);\s+/\* 0x([0-9a-f]+) \*/", l)
    if m:
        m2 = re.match(r"\s+/\* 0x([0-9a-f]+) \*/", txt[i + 1])
        ins.append((int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16), int(m2.group(1), 16)))
targets = set()
for a, t, w1, w2 in ins:
    m = re.search(r"BRA(?:\.\S+)? .*?(0x[0-9a-f]+)", t)
    if m: targets.add(int(m.group(1), 16))
blocks, cur = [], []
for a, t, w1, w2 in ins:
    if a in targets and cur: blocks.append(cur); cur = []
    cur.append((a, t, w2))
    if re.match(r"(@!?U?P\d+\s+)?(BRA|EXIT|RET|BSYNC)", t): blocks.append(cur); cur = []
if cur: blocks.append(cur)
for b in blocks:
    if len(b) < min_n: continue
    st = sum((w2 >> 41) & 0xF for _, _, w2 in b)
    ops = collections.Counter()
    sto = collections.Counter()
    for _, t, w2 in b:
        op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]
        ops[op] += 1; sto[op] += (w2 >> 41) & 0xF
    waits = sum(1 for _, _, w2 in b if (w2 >> 52) & 0x3F)
    print(f"block {b[0][0]:#x}..{b[-1][0]:#x}: {len(b)} instr, sum(stall) {st}, {waits} instr wait on a scoreboard")
    print("   " + "  ".join(f"{o}:{n}/{sto[o]}" for o, n in ops.most_common(14)))

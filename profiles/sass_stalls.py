"""Static schedule of the loops of a kernel: for every backward branch of a cubin/object, the
instructions in the loop body and the sum of the stall counts ptxas encoded (a lower bound of the
cycles one warp needs per trip, before scoreboard waits).
usage: cuobjdump -sass x.o > x.sass; python profiles/sass_stalls.py x.sass [mnemonic filter, e.g. FMNMX]"""
import re
import sys

lines = open(sys.argv[1]).read().split("\n")
flt = sys.argv[2] if len(sys.argv) > 2 else None
ins = []
i = 0
while i < len(lines):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/", lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r"\s+/\* (0x[0-9a-f]{16}) \*/", lines[i + 1])
        if m2:
            hi = int(m2.group(1), 16)
            ins.append((int(m.group(1), 16), m.group(2).strip(), (hi >> 41) & 0xf))
            i += 2
            continue
    if "Function :" in lines[i]:
        ins.append((-1, lines[i].strip(), 0))
    i += 1
addr = {}
for k, (a, t, s) in enumerate(ins):
    if a >= 0:
        addr.setdefault(a, []).append(k)
for k, (a, t, s) in enumerate(ins):
    m = re.search(r"BRA\S* (?:\S+, )?0x([0-9a-f]+)$", t)
    if not m or a < 0:
        continue
    tgt = int(m.group(1), 16)
    if tgt >= a:
        continue
    # the target instruction of the same function = the closest earlier index with that address
    cands = [j for j in addr.get(tgt, []) if j < k]
    if not cands:
        continue
    j = cands[-1]
    body = ins[j:k + 1]
    if flt and not any(flt in b[1] for b in body):
        continue
    n = len(body)
    if n > 700:
        continue
    st = sum(b[2] for b in body)
    cnt = {}
    for b in body:
        op = b[1].split()[0] if not b[1].startswith("@") else b[1].split()[1]
        op = op.split(".")[0]
        cnt[op] = cnt.get(op, 0) + 1
    top = ", ".join(f"{o}:{c}" for o, c in sorted(cnt.items(), key=lambda kv: -kv[1])[:8])
    print(f"loop {tgt:05x}..{a:05x}: {n:4d} instr, stall sum {st:5d}  [{top}]")

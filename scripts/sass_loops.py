"""Instruction mix of the loops of one device function inside a cuobjdump -sass listing.
usage: sass_loops.py listing.sass start_hex size_hex"""
import bisect
import re
import sys

lines = open(sys.argv[1]).read().split("\n")
lo0, hi0 = int(sys.argv[2], 16), int(sys.argv[2], 16) + int(sys.argv[3], 16)
ins = []
for l in lines:
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m and lo0 <= int(m.group(1), 16) < hi0:
        ins.append((int(m.group(1), 16), m.group(2)))
print(len(ins), "instructions")
addrs = [a for a, _ in ins]
loops = []
for a, t in ins:
    m = re.search(r"BRA\S*\s+.*0x([0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if lo0 <= tgt < a:
            loops.append((tgt, a))


def stats(lo, hi):
    seg = [t for _, t in ins[bisect.bisect_left(addrs, lo):bisect.bisect_right(addrs, hi)]]
    c = lambda k: sum(1 for t in seg if re.search(k, t))
    return dict(n=len(seg), dfma=c("DFMA"), dmul=c("DMUL"), dadd=c("DADD"), bar=c("BAR"), ldl=c("LDL"), stl=c("STL"),
                lds=c(r"\bLDS"), sts=c(r"\bSTS"), shfl=c("SHFL"), ld=c(r"\bLD\."), st=c(r"\bST\."), ldg=c("LDG"), stg=c("STG"),
                imad=c("IMAD"), call=c("CALL"), bra=c("BRA"), mufu=c("MUFU"))


for lo, hi in sorted(loops, key=lambda x: x[0] - x[1])[:int(sys.argv[4]) if len(sys.argv) > 4 else 5]:
    print(hex(lo), hex(hi), hi - lo, stats(lo, hi))

"""List the loops (backward branches) of a cuobjdump -sass dump with their instruction mix.
usage: cuobjdump -sass -fun NAME lib.so | python tools/sass_loops.py [min_instructions]"""
import re, sys, collections
mn = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ins = []
for line in sys.stdin:
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
addr2i = {a: i for i, (a, _) in enumerate(ins)}
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr2i and i - addr2i[tgt] >= mn:
            body = ins[addr2i[tgt]:i + 1]
            c = collections.Counter()
            for _, s in body:
                s = re.sub(r"^@!?U?P\d+\s+", "", s)
                c[s.split()[0].split(".")[0]] += 1
            print("loop 0x%x..0x%x  %d instr: %s" % (tgt, a, len(body), " ".join("%s=%d" % kv for kv in c.most_common(14))))

"""Instruction mix of the hot loops of a kernel: python scripts/sass_loops.py <mangled kernel name> [marker]"""
import collections, re, subprocess, sys
fun, marker = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
lib = sys.argv[3] if len(sys.argv) > 3 else "anovos_b200/libanovos_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", "-fun", fun, lib], capture_output=True, text=True).stdout
lines = [re.sub(r"/\* 0x[0-9a-f]* \*/", "", l) for l in txt.split("\n") if re.match(r"^\s+/\*[0-9a-f]{4,6}\*/", l)]
addr = lambda l: int(re.search(r"/\*([0-9a-f]+)\*/", l).group(1), 16)
amap = {addr(l): i for i, l in enumerate(lines)}
for i, l in enumerate(lines):
    m = re.search(r"BRA\S*\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", l)
    if not m:
        continue
    t, a = int(m.group(1), 16), addr(l)
    if t < a and t in amap:
        body = lines[amap[t]:i + 1]
        s = "\n".join(body)
        n128 = s.count(".128")
        if n128 >= 4 and len(body) < 4000 and (marker in s):
            ops = []
            for b in body:
                p = re.sub(r"\s+", " ", b.split("*/")[1]).strip().split(" ")
                ops.append(p[1] if p[0].startswith("@") else p[0])
            c = collections.Counter(ops)
            per = 16 // (8 if ("F2F" not in s and "DADD" in s and "I2F" not in s and False) else 4)
            nelem = n128 * 4
            print("loop @%x len=%d LDG128=%d bitmap_loads=%d -> %.1f instr per 4-byte element" % (t, len(body), n128, s.count("LDG.E.CONSTANT"), len(body) / nelem))
            print("   ", {k: round(v / nelem, 2) for k, v in c.most_common(28)})

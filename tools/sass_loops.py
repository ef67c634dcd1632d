"""List the loops of a kernel's SASS (cuobjdump -sass output on stdin or a file): address range,
instruction count, local-memory operations and a histogram of the busiest mnemonics.

    cuobjdump -sass -fun <mangled> file.o | python tools/sass_loops.py [--min 100]
"""
import collections
import re
import sys


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    min_len = 100
    if "--min" in sys.argv:
        min_len = int(sys.argv[sys.argv.index("--min") + 1])
        args = [a for a in args if a != str(min_len)]
    text = open(args[0]).read() if args else sys.stdin.read()
    ins = []
    for m in re.finditer(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", text):
        ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.\w+)*\s+(?:[!\w]+,\s*)?`?\(?\.?L?_?x?_?\w*\)?", t)
        m2 = re.search(r"BRA.*?0x([0-9a-f]+)", t)
        if t.split()[0].startswith("@"):
            op = t.split()[1]
        else:
            op = t.split()[0]
        if op.startswith("BRA") and m2:
            target = int(m2.group(1), 16)
            if target <= a and target in addr_index:
                j = addr_index[target]
                if i - j + 1 >= min_len:
                    loops.append((j, i))
    for j, i in loops:
        body = [t for _, t in ins[j:i + 1]]
        ops = collections.Counter()
        for t in body:
            parts = t.split()
            op = parts[1] if parts[0].startswith("@") else parts[0]
            ops[op.split(".")[0]] += 1
        local = sum(v for k, v in ops.items() if k in ("LDL", "STL"))
        top = ", ".join("%s %d" % kv for kv in ops.most_common(10))
        print("loop 0x%05x-0x%05x  %4d instr  local %d  | %s" % (ins[j][0], ins[i][0], i - j + 1, local, top))


if __name__ == "__main__":
    main()

"""Per-kernel SASS summary of libfovea_b200.so (runs without a GPU): instruction count and the memory / synchronisation /
conversion mnemonics that characterise each kernel.  `python tools/sass_summary.py > profiles/r2_sass_summary.txt`"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "foveated-instance-segmentation_b200", "fovea", "libfovea_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda names: subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
KEEP = re.compile(r"^(LD|ST|ATOM|RED|BAR|SHFL|MATCH|REDUX|MUFU|UTMA|UBLKCP|SYNCS|TCGEN|UTC|LDGSTS|I2F|F2F|D(ADD|MUL|FMA)|VOTE|CCTL|MEMBAR|FENCE|ELECT)")
kern, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); kern[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        kern[cur]["__n"] += 1
        op = m.group(1)
        if KEEP.match(op):
            kern[cur][op] += 1
names = demangle(list(kern))
print("SASS of foveated-instance-segmentation_b200/fovea/libfovea_b200.so (cuobjdump -sass, sm_100a): per kernel, the instruction count and")
print("the memory / synchronisation mnemonics that characterise it.  Regenerate: python tools/sass_summary.py\n")
for (k, c), name in zip(kern.items(), names):
    name = re.sub(r"\(.*", "", name)
    print(f"{name}: {c['__n']} instructions")
    print("    " + ", ".join(f"{op} x{n}" for op, n in sorted(c.items()) if op != "__n"))

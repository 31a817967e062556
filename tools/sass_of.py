"""Dump the SASS of one kernel of libb200vis.so (substring match on the mangled name) with opcode histogram.
   python tools/sass_of.py k_propagate_cull_leanILb1ELb1ELb1E [out.txt]"""
import subprocess, sys, re, collections
pat = sys.argv[1]
out = subprocess.run(["cuobjdump", "-sass", "bevy_b200/libb200vis.so"], capture_output=True, text=True).stdout
cur = None; lines = []
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m: cur = m.group(1); continue
    if cur and pat in cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
        lines.append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", ln).rstrip())
print(len(lines), "instructions")
ops = collections.Counter()
for ln in lines:
    t = ln.split("*/", 1)[1].split()
    op = t[1] if t[0].startswith("@") else t[0]
    ops[op.split(".")[0]] += 1
print(ops.most_common(25))
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write("\n".join(f"{i:5d} {l.split('*/',1)[1].strip()}" for i, l in enumerate(lines)) + "\n")

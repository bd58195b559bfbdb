"""Instructions of one kernel whose OUTERMOST inlining frame is source line <outer>, broken down by innermost line.
usage: python tools/sass_inner.py <cubin> <kernel substring> <source path> <outer line> [min count]"""
import collections
import re
import subprocess
import sys

cubin, kern, srcpath, outer_line = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
minc = int(sys.argv[5]) if len(sys.argv) > 5 else 6
txt = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout
src = open(srcpath).read().splitlines()
base = srcpath.split("/")[-1]
inside, pend, cur, cnt = False, [], None, collections.Counter()
for line in txt.splitlines():
    if line.startswith("\t.section\t.text."):
        inside = kern in line
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
    if m:
        pend.append((m.group(1).split("/")[-1], int(m.group(2)), int(m.group(4)) if m.group(4) else None))
        continue
    if re.match(r"^\s*/\*[0-9a-f]{4,}\*/", line):
        if pend:
            cur, pend = pend, []
        if cur:
            o = cur[-1]
            if (o[2] if o[2] else o[1]) == outer_line:
                cnt[(cur[0][0], cur[0][1])] += 1
tot = 0
for (f, l), v in sorted(cnt.items()):
    tot += v
    if v >= minc:
        print(f, l, v, src[l - 1].strip()[:100] if f == base else "")
print("total", tot)

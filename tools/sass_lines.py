"""Attribute the SASS instructions of one kernel to source lines (outermost frame of the inlining chain).
usage: python tools/sass_lines.py <cubin> <kernel name substring> <source file basename> [lo hi]"""
import collections
import re
import subprocess
import sys


def main():
    cubin, kern, src = sys.argv[1:4]
    lo, hi = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (0, 1 << 30)
    txt = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout
    cnt = collections.Counter()
    cur = None
    total = 0
    pend = []
    inside = False
    for line in txt.splitlines():
        if line.startswith("\t.section\t.text."):
            inside = kern in line
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
        if m:
            f, l = (m.group(3), m.group(4)) if m.group(3) else (m.group(1), m.group(2))
            pend.append((f.split("/")[-1], int(l)))
            continue
        if re.match(r"^\s*/\*[0-9a-f]{4,}\*/", line):
            if pend:
                own = [p for p in pend if p[0] == src]
                cur = own[-1] if own else pend[-1]
                pend = []
            total += 1
            if cur:
                cnt[cur] += 1
    print("total", total)
    for (f, l), v in sorted(cnt.items()):
        if f == src and lo <= l <= hi:
            print(f"{l:5d} {v}")


if __name__ == "__main__":
    main()

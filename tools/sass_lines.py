#!/usr/bin/env python
"""SASS instructions per source line of one kernel (static code size, from `nvdisasm -g -c` of a cubin built with -lineinfo).
    cuobjdump -xelf all file.o && python tools/sass_lines.py file.sm_100a.cubin <substring of the kernel's mangled name> [top]
Code size matters on this path: a long straight-line body that does not fit the instruction caches is re-fetched every frame pair."""
import collections
import re
import subprocess
import sys


def main():
    cubin, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    cur, on, cnt = None, False, collections.Counter()
    for line in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
        if m:
            on = pat in m.group(1)
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line) and cur:
            cnt[cur] += 1
    tot = sum(cnt.values())
    print("total SASS instructions: %d (%.1f KB)" % (tot, tot * 16 / 1024))
    for (f, ln), v in cnt.most_common(top):
        print("%6d  %s:%d" % (v, f, ln))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Static SASS size of a kernel, attributed to source regions (nvdisasm -g line info)."""
import collections
import re
import sys

txt = open(sys.argv[1]).read().split("\n")
for kn in sys.argv[2:]:
    s = [i for i, l in enumerate(txt) if l.startswith(".text.") and kn in l][0]
    e = [i for i, l in enumerate(txt) if i > s and l.startswith("//---------------------")]
    e = e[0] if e else len(txt)
    cnt, cf, cl, n = collections.Counter(), None, None, 0
    for l in txt[s:e]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cf, cl = m.group(1).split("/")[-1], int(m.group(2))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+.*?;", l):
            cnt[(cf, cl // 25 * 25)] += 1
            n += 1
    print(kn, "static SASS instr", n, "=", n * 16 / 1024, "KB")
    for (f, l), c in sorted(cnt.items(), key=lambda kv: -kv[1])[:40]:
        print(f"   {c:6d} ({c * 16 / 1024:5.1f} KB)  {f}:{l}-{l + 24}")

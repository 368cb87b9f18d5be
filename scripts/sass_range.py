#!/usr/bin/env python
"""sass_range.py <file.sass> <function-substring> <start-hex> <end-hex> — print the SASS of one address range."""
import re, sys
fn = [f for f in open(sys.argv[1]).read().split("Function : ")[1:] if sys.argv[2] in f.split("\n", 1)[0]][0]
lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
for ln in fn.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and lo <= int(m.group(1), 16) <= hi:
        print(f"{m.group(1)}  {m.group(2).strip()}")

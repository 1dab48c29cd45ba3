#!/usr/bin/env python3
"""Static code size of one kernel per source region: nvdisasm -g -c <cubin> | tools/sass_regions.py <kernel substring>
Counts SASS instructions per (file, line) from the `//## File "...", line N` annotations (innermost inlined location) and
prints the lines / files with the most instructions — what the instruction cache has to hold."""
import re, sys, collections
kern = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; loc = None; insec = False
per_line = collections.Counter(); per_file = collections.Counter(); total = 0
for ln in sys.stdin:
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        insec = kern in m.group(1); continue
    if ln.startswith("//---") : 
        continue
    if not insec: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        loc = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln) and loc:
        per_line[loc] += 1; per_file[loc[0]] += 1; total += 1
print("kernel *%s*: %d SASS instructions (%.1f KB)" % (kern, total, total * 16 / 1024))
for f, c in per_file.most_common(): print("  %-28s %5d" % (f, c))
print("  -- top lines")
for (f, l), c in per_line.most_common(top): print("  %-28s %5d  %4.1f%%" % ("%s:%d" % (f, l), c, 100 * c / max(total, 1)))

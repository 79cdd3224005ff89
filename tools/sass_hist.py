"""Per-function SASS opcode histogram of an object / shared library:  python tools/sass_hist.py <file> [name filter]"""
import collections, re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
flt = sys.argv[2] if len(sys.argv) > 2 else ""
cnt, ops, name = collections.Counter(), collections.defaultdict(collections.Counter), None
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = m.group(1)
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)', line)
    if m and name and flt in name:
        cnt[name] += 1
        ops[name][m.group(1).split('.')[0]] += 1
for n, c in cnt.most_common():
    print(c, n[:90], dict(ops[n].most_common(10)))

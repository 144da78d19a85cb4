"""dev tool: per-iteration instruction mix of a kernel from an object file: sassmix.py <obj> <mangled-name>"""
import subprocess, sys
from collections import Counter
obj, fun = sys.argv[1], sys.argv[2]
txt = subprocess.run(['cuobjdump', '-sass', '-fun', fun, obj], capture_output=True, text=True).stdout
ops = []
for line in txt.splitlines():
    line = line.strip()
    if line.startswith('/*') and '*/' in line:
        rest = line.split('*/', 1)[1].strip()
        if rest and rest[0].isalpha() or rest.startswith('@'):
            tok = rest.split()
            op = tok[1] if tok[0].startswith('@') else tok[0]
            ops.append(op.rstrip(';'))
base = [o.split('.')[0] for o in ops]
print('total', len(ops))
print(Counter(base).most_common(14))
print('full', Counter(o for o in ops if o.split('.')[0] in ('LDG','STG','LDS','STS','UBLKCP','SYNCS')).most_common())

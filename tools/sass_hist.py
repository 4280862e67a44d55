"""Opcode histogram of one kernel's SASS (static counts).  usage: python tools/sass_hist.py lib.so [kernel-substring]"""
import collections, os, re, subprocess, sys, tempfile
so = os.path.abspath(sys.argv[1]); key = sys.argv[2] if len(sys.argv) > 2 else 'pbs_kernel'
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', so], cwd=tmp, capture_output=True)
cubin = max((f for f in os.listdir(tmp) if f.endswith('.cubin')), key=lambda f: os.path.getsize(os.path.join(tmp, f)))
txt = subprocess.run(['nvdisasm', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
m = re.search(r'\.text\.(\S*%s\S*):' % key, txt)
start = m.start(); end = txt.find('//--------------------- .', start + 10)
c = collections.Counter()
for l in txt[start:end].splitlines():
    mm = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if mm:
        t = mm.group(2).split()
        op = t[1] if t[0].startswith('@') else t[0]
        c[op.split('.')[0]] += 1
tot = sum(c.values())
print(m.group(1), 'total', tot)
print(' '.join(f'{k}:{v}' for k, v in c.most_common(40)))

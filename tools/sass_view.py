"""One character per SASS instruction of a kernel (F fp64, l LDS, s STS, G LDG, | BAR, t/T tensor-memory ld/st, L/S local ld/st):
shows at a glance how loads, stores and barriers interleave with the arithmetic.  usage: python tools/sass_view.py lib.so [kernel]"""
import os, re, subprocess, sys, tempfile
so = os.path.abspath(sys.argv[1]); key = sys.argv[2] if len(sys.argv) > 2 else 'pbs_kernel'
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', so], cwd=tmp, capture_output=True)
cubin = max((f for f in os.listdir(tmp) if f.endswith('.cubin')), key=lambda f: os.path.getsize(os.path.join(tmp, f)))
txt = subprocess.run(['nvdisasm', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
m = re.search(r'\.text\.(\S*%s\S*):' % key, txt)
start = m.start(); end = txt.find('//--------------------- .', start + 10)
cls = {'DFMA': 'F', 'DADD': 'F', 'DMUL': 'F', 'LDS': 'l', 'STS': 's', 'LDG': 'G', 'BAR': '|', 'LDTM': 't', 'STTM': 'T', 'LDL': 'L', 'STL': 'S',
       'SYNCS': 'm', 'UBLKCP': 'B'}
s = ''
for l in txt[start:end].splitlines():
    mm = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if mm:
        t = mm.group(2).split(); op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        s += cls.get(op, '.')
for i in range(0, len(s), 160):
    print(f'{i:5d} {s[i:i + 160]}')

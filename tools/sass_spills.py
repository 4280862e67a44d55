"""Where a kernel spills: STL / LDL instructions per source line.  usage: python tools/sass_spills.py lib.so [kernel-substring]"""
import collections, os, re, subprocess, sys, tempfile
so = os.path.abspath(sys.argv[1]); key = sys.argv[2] if len(sys.argv) > 2 else 'pbs_kernel'
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', so], cwd=tmp, capture_output=True)
cubin = max((f for f in os.listdir(tmp) if f.endswith('.cubin')), key=lambda f: os.path.getsize(os.path.join(tmp, f)))
txt = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
m = re.search(r'\.text\.(\S*%s\S*):' % key, txt)
start = m.start(); end = txt.find('//--------------------- .', start + 10)
cur = None; c = collections.Counter(); n = 0; pos = {}
for l in txt[start:end].splitlines():
    mm = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if mm: cur = (os.path.basename(mm.group(1)), int(mm.group(2))); continue
    mm = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if mm:
        n += 1
        t = mm.group(2).split(); op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        if op in ('STL', 'LDL'):
            c[(cur, op)] += 1; pos.setdefault((cur, op), n)
for (loc, op), v in sorted(c.items(), key=lambda kv: pos[kv[0]]):
    print(f'{pos[(loc, op)]:6d} {op} x{v} {loc}')
print('instructions', n)

"""Per-barrier-segment opcode histogram of a kernel's SASS (where did the compiler put the loads?).
usage: python tools/sass_segments.py lib.so [kernel-substring]"""
import collections, os, re, subprocess, sys, tempfile
so = os.path.abspath(sys.argv[1]); key = sys.argv[2] if len(sys.argv) > 2 else 'pbs_kernel'
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', so], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
txt = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
m = re.search(r'\.text\.(\S*%s\S*):' % key, txt)
start = m.start(); end = txt.find('//--------------------- .', start + 10)
ins, cur = [], None
for l in txt[start:end].splitlines():
    mm = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if mm: cur = (os.path.basename(mm.group(1)), int(mm.group(2))); continue
    mm = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if mm:
        t = mm.group(2).split()
        op = t[1] if t[0].startswith('@') else t[0]
        ins.append((op.split('.')[0], cur))
bars = [i for i, (o, c) in enumerate(ins) if o == 'BAR']
prev = 0
for b in bars + [len(ins)]:
    seg = ins[prev:b]
    c = collections.Counter(o for o, _ in seg)
    keep = ['DADD', 'DMUL', 'DFMA', 'LDS', 'STS', 'LDG', 'STL', 'LDL']
    other = sum(v for k, v in c.items() if k not in keep)
    print(f"{prev:5d}-{b:5d} n={b-prev:4d} " + ' '.join(f"{k}:{c[k]}" for k in keep if c[k]) + f" other:{other}")
    prev = b + 1

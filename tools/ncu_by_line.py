"""Join an ncu per-SASS-instruction profile with nvdisasm line info (same build!) and aggregate
stall samples / executed instructions per source line and per coarse phase.
usage: python tools/ncu_by_line.py prof.ncu-rep spf_b200/libspf_b200.so <mangled kernel name> [out.csv]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(so, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(so)], cwd=tmp, capture_output=True)
    cubin = max((f for f in os.listdir(tmp) if f.endswith('.cubin')), key=lambda f: os.path.getsize(os.path.join(tmp, f)))
    txt = subprocess.run(['nvdisasm', '-g', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    start = txt.index('.text.' + kernel + ':')
    end = txt.find('//--------------------- .', start)
    cur, out = ('?', 0), []
    for line in txt[start:end].splitlines():
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', line):
            out.append(cur)
    return out


def main():
    rep, so, kernel = sys.argv[1:4]
    locs = sass_lines(so, kernel)
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    ins = [r for r in rows[2:] if len(r) >= len(hdr)]
    if len(ins) != len(locs):
        print(f'WARNING: {len(ins)} profiled instructions vs {len(locs)} in the binary: not the same build', file=sys.stderr)
    per_line = collections.defaultdict(lambda: [0, 0])
    for r, loc in zip(ins, locs):
        per_line[loc][0] += int(r[idx['# Samples']] or 0)
        per_line[loc][1] += int(r[idx['Instructions Executed']] or 0)
    tot = sum(v[0] for v in per_line.values()) or 1
    srcs = {}
    out = ['file,line,samples_pct,exec_M,source']
    for (f, l), (s, e) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:60]:
        if f not in srcs:
            p = os.path.join('spf_b200', 'csrc', f)
            srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = srcs[f][l - 1].strip()[:80].replace(',', ';') if 0 < l <= len(srcs[f]) else ''
        out.append(f'{f},{l},{100 * s / tot:.2f},{e / 1e6:.1f},{text}')
    text = '\n'.join(out) + '\n'
    if len(sys.argv) > 4:
        open(sys.argv[4], 'w').write(text)
    print(text)


if __name__ == '__main__':
    main()

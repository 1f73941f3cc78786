"""Attribute ncu warp-stall samples / executed instructions of one kernel to CUDA source lines.
usage: python profiles/attribute_lines.py <report.ncu-rep> <kernel-name-substring> <nvdisasm -g output> [top_n]"""
import collections, csv, re, subprocess, sys
rep, kern, sassf = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 24
sass = open(sassf).read().split('\n')
start = [i for i, l in enumerate(sass) if l.startswith('.text.') and kern in l][0]
end = [i for i, l in enumerate(sass) if i > start and l.startswith('//--------------------- .text.')]
end = end[0] if end else len(sass)
cur, seq = None, []
for l in sass[start:end]:
    m = re.search(r'//## File "(.*)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4}\*/', l):
        seq.append(cur)
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{kern}'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]
si, ie = hdr.index('# Samples'), hdr.index('Instructions Executed')
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[si].isdigit()][:len(seq)]
bl, bs = collections.Counter(), collections.Counter()
for k, r in enumerate(data):
    bl[seq[k]] += int(r[ie] or 0)
    bs[seq[k]] += int(r[si])
toti, tots = sum(bl.values()), sum(bs.values())
print(f'{kern}: {len(seq)} SASS instructions, {toti} warp instructions executed, {tots} samples')
files = {}
for (key, c) in sorted(bs.items(), key=lambda x: -x[1])[:top]:
    if key is None:
        continue
    f, line = key
    if f not in files:
        try:
            files[f] = open(f).read().split('\n')
        except OSError:
            files[f] = []
    text = files[f][line - 1].strip()[:95] if line - 1 < len(files[f]) else ''
    print(f'{100 * c / tots:5.1f}% samp {100 * bl[key] / toti:5.1f}% instr  L{line}: {text}')

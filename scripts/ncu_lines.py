"""Attribute executed warp-instructions of one kernel (ncu source page, SASS) to CUDA source lines using nvdisasm --print-line-info of the
matching cubin.  usage: ncu_lines.py rep.ncu-rep disasm.txt mangled_kernel_name [topN]"""
import csv, io, re, subprocess, sys
from collections import Counter
rep, dis, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; iA, iE, iS = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
ex = []
for r in rows[2:]:
    if r and r[0] == 'Kernel Name': break
    if len(r) < len(hdr) or not r[0].startswith('0x'): continue
    ex.append((r[iA].strip(), int(r[iE]), int(r[iS])))
lines = open(dis).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.' + kern + ':'))
cur = None; k = 0; agg = Counter(); smp = Counter(); inl = {}
for l in lines[start + 1:]:
    if l.startswith('.text.') or l.startswith('.section'): break
    m = re.search(r'//## File "([^"]*)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', l):
        if k < len(ex):
            agg[cur] += ex[k][1]; smp[cur] += ex[k][2]
        k += 1
print("sass instructions: ncu %d, disasm %d; total executed %.1f M" % (len(ex), k, sum(e[1] for e in ex) / 1e6))
srcs = {}
for (f, ln), e in agg.most_common(top):
    if f not in srcs:
        try: srcs[f] = open('/root/repo/allscale_ipic3d_b200/csrc/' + f).read().split('\n')
        except Exception: srcs[f] = []
    text = srcs[f][ln - 1].strip()[:100] if ln - 1 < len(srcs[f]) else ''
    print("%8.2f M  smp %6d  %s:%d  %s" % (e / 1e6, smp[(f, ln)], f, ln, text))

"""Per-kernel SASS evidence for profiles/: which Blackwell-era instructions each kernel of libipic3d_b200.so contains
(UBLKCP = cp.async.bulk, the TMA engine's 1-D bulk copy; SYNCS = mbarrier operations; MATCH = match.any; DFMA/DMUL/DADD = FP64
pipe; MUFU.RSQ64H / RCP64H = FP64 reciprocal seeds; STG...EF = evict-first streaming stores), with the first occurrences quoted.
usage: python scripts/sass_excerpt.py > profiles/<round>_sass_excerpt.txt   (needs cuobjdump; no GPU)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "allscale_ipic3d_b200", "lib", "libipic3d_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
kernels = collections.OrderedDict()
cur = None
for line in sass.split("\n"):
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); kernels[cur] = []
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        kernels[cur].append(line.strip())
KEYS = ["UBLKCP", "SYNCS", "MATCH", "UCGABAR", "CGAERRBAR", "DFMA", "DMUL", "DADD", "MUFU.RSQ64H", "MUFU.RCP64H", "BAR.SYNC", "SHFL", "VOTE", "ATOM", "RED", "LDS", "STS", "LDG", "STG"]
print("SASS of %s (sm_100a), instructions per kernel; arch check: %s" % (os.path.basename(lib), re.search(r"arch = (\S+)", sass).group(1)))
for name, ins in kernels.items():
    short = demangle(name)
    m = re.match(r"^(?:void )?(ipic::\w+(?:<.*?>)?)\(", short)
    short = re.sub(r"\((int|bool)\)", "", m.group(1)) if m else short
    if not re.search(r"k_move|k_import|k_sort|k_big|k_face|k_init|k_solver_wavefront|k_solver_cluster|k_gather_current|k_trace", short):
        continue
    c = collections.Counter()
    first = {}
    for i in ins:
        op = re.sub(r"^/\*[0-9a-f]+\*/\s+", "", i)
        op = re.sub(r"^@!?U?P\d+\s+", "", op)
        for k in KEYS:
            if op.startswith(k):
                c[k] += 1
                first.setdefault(k, op.split(";")[0])
    print("\n%s  [%d SASS instructions]" % (short, len(ins)))
    print("   " + "  ".join("%s %d" % (k, c[k]) for k in KEYS if c[k]))
    for k in ("UBLKCP", "SYNCS", "MATCH", "UCGABAR"):
        if k in first:
            print("   e.g. " + first[k])

"""dev/similarity_check.py -- self-check: similarity of every source file of this repo to the SAME-NAMED file of the
reference tree (line-based and comment-stripped token-based difflib ratios).  Needs /root/reference."""
import difflib, os, re, sys
ref = '/root/reference'; mine = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def norm(t):
    t2 = re.sub(r'//.*', '', t); t2 = re.sub(r'/\*.*?\*/', '', t2, flags=re.S)
    return re.sub(r'\s+', ' ', t2)
res = []
for root, _, files in os.walk(mine):
    if any(x in root for x in ('.git', '_build', '_ref', 'gpurun_out', '/build')): continue
    for f in files:
        if not f.endswith(('.cuh', '.cu', '.cpp', '.h', '.hpp', '.py', '.inc')): continue
        p = os.path.join(root, f)
        for rroot, _, rfiles in os.walk(ref):
            if f in rfiles:
                a = open(p, errors='ignore').read(); b = open(os.path.join(rroot, f), errors='ignore').read()
                if difflib.SequenceMatcher(None, a, b, autojunk=False).quick_ratio() > 0.4:
                    r = difflib.SequenceMatcher(None, a.splitlines(), b.splitlines(), autojunk=False).ratio()
                    rn = difflib.SequenceMatcher(None, norm(a).split(' '), norm(b).split(' '), autojunk=False).ratio()
                    rc = difflib.SequenceMatcher(None, norm(a), norm(b), autojunk=False).ratio() if len(a) < 20000 else -1
                    res.append((max(r, rn), r, rn, rc, p.replace(mine + '/', '')))
res.sort(reverse=True)
for x in res[:int(sys.argv[1]) if len(sys.argv) > 1 else 20]: print("max=%.2f line=%.2f token=%.2f char=%.2f %s" % x)

"""SASS listings + mnemonic counts of the hot kernels from the shipped library:  python profiles/extract_sass.py
Writes profiles/r02_sass/<kernel>.sass and SUMMARY.txt (cuobjdump -sass, sm_100a)."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multiple-object-tracking-lidar_b200", "libmot_b200.so")
OUT = os.path.join(ROOT, "profiles", "r02_sass")
WANT = ["k_cell_localE", "k_cell_local_dense", "k_compact_onepassILi0", "k_rs_scatterIjLb0ELi16ELi512", "k_rs_scatterIjLb0ELi8ELi256", "k_uf_fusedIj", "k_uf_heavy",
        "k_uf_sparse2", "k_fs_front", "k_fs_edges", "k_fs_tables"]
GROUPS = [("UBLKCP", r"^UBLKCP"), ("SYNCS", r"^SYNCS"), ("UCGABAR", r"^UCGABAR_ARV"), ("ATOMG", r"^ATOMG|^ATOM\."), ("ATOMS", r"^ATOMS"), ("RED", r"^RED"),
          ("SHFL", r"^SHFL"), ("VOTE", r"^VOTE"), ("MATCH", r"^MATCH"), ("REDUX", r"^REDUX"), ("LDG", r"^LDG"), ("LD", r"^LD\."), ("STG", r"^STG"), ("LDS", r"^LDS"),
          ("STS", r"^STS"), ("FADD", r"^FADD"), ("FMUL", r"^FMUL"), ("FFMA", r"^FFMA"), ("FMNMX", r"^FMNMX"), ("DADD", r"^DADD"), ("BAR", r"^BAR"),
          ("BSSY", r"^BSSY"), ("WARPSYNC", r"^WARPSYNC"), ("LDL", r"^LDL"), ("STL", r"^STL")]
text = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs = collections.OrderedDict()
cur = None
for line in text.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur:
        funcs[cur].append(line)
os.makedirs(OUT, exist_ok=True)
for f in os.listdir(OUT):
    if f.endswith(".sass"):
        os.remove(os.path.join(OUT, f))
rows = []
for w in WANT:
    names = [n for n in funcs if w in n]
    if not names:
        continue
    n = names[0]
    body = funcs[n]
    with open(os.path.join(OUT, w + ".sass"), "w") as fo:
        fo.write("Function : " + n + "\n" + "\n".join(body) + "\n")
    ops = []
    for line in body:
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            ops.append(m.group(1))
    cnt = [(g, sum(1 for o in ops if re.match(rx, o))) for g, rx in GROUPS]
    rows.append(f"{w + '.sass':36s}{len(ops):5d} instr  " + " ".join(f"{g}={c}" for g, c in cnt if c))
head = """# SASS mnemonic counts of the hot kernels (cuobjdump -sass of the shipped libmot_b200.so, sm_100a); listings beside this file
# (regenerate: python profiles/extract_sass.py).
# UBLKCP = cp.async.bulk (TMA 1-D bulk copy), SYNCS = mbarrier, UCGABAR = thread-block-cluster barrier, ATOMG/ATOMS/RED = atomics,
# SHFL/VOTE/MATCH/REDUX = warp collectives, LD = generic loads (distributed shared memory / the argument block of the small-frame kernels).
# No FFMA in the union-find kernels (built with -fmad=false; the distance predicate is FADD / FMUL only); the FFMA of k_compact_onepass<map>
# and k_fs_front are the expansion of the two IEEE divisions (__fdiv_rn) of removeStatic's index arithmetic, not contractions.
"""
open(os.path.join(OUT, "SUMMARY.txt"), "w").write(head + "\n".join(rows) + "\n")
print(head + "\n".join(rows))

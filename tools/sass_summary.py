"""Opcode histogram of the hot kernels in the shipped library (cuobjdump -sass), the evidence for
packed math (FFMA2/FADD2/FMUL2), wide accesses (LDG.E.64 / STG.E.64, LDG.E.ENL2.256 / STG.E.ENL2.256), L2 bulk prefetch (UBLKPF) and TMA
bulk copies (UBLKCP).   python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fvdbm_jax_b200", "libfvdbm_b200.so")
KERNELS = ["k_fused_recILi3ELi1E", "k_fused_recILi3ELi0E", "k_fused_pairILi9ELi3ELi1E", "k_fused_pairILi9ELi3ELi0E", "k_fused_directIfLi9ELi3ELi1E", "k_fused_directIdLi9ELi3ELi1E",
           "k_fused_tmaIfLi9ELi3ELi1E", "k_nodesIfLi9E"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", out)
head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
print(f"# cuobjdump -sass fvdbm_jax_b200/libfvdbm_b200.so (sm_100a), built from the tree at/after commit {head}")
print("# static instruction counts per kernel (all paths, incl. the rare ghost-side branches)\n")
for k in KERNELS:
    for b in blocks[1:]:
        name = b.split("\n", 1)[0].strip()
        if k in name:
            ops = collections.Counter()
            for m in re.finditer(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", b, re.M):
                op = m.group(1)
                base = op.split(".")[0]
                if base in ("LDG", "STG", "LDS", "STS"):
                    w = re.search(r"\.(64|128|256)", op)
                    base += "." + (w.group(1) if w else "32")
                ops[base] += 1
            tot = sum(ops.values())
            print(f"{name}\n  total {tot}: " + ", ".join(f"{o} {c}" for o, c in ops.most_common(28)))
            keys = ["FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "FSEL", "DFMA", "DADD", "DMUL", "LDG.32", "LDG.64", "LDG.128", "LDG.256", "STG.32",
                    "STG.64", "STG.128", "STG.256", "UBLKPF", "UBLKCP", "SYNCS", "SHFL"]
            print("  key: " + ", ".join(f"{x}={ops.get(x, 0)}" for x in keys) + "\n")
            break

"""SASS evidence for the built (git-ignored) libslcl.so: per kernel, how many of the instructions that prove the
Blackwell paths are really in the binary.

    python tools/sass_summary.py > profiles/sass_summary.txt

UTCHMMA = tcgen05.mma (kind::f16), LDTM/STTM = tcgen05.ld/st (tensor memory), UTMALDG/UTMASTG = TMA tensor loads/stores,
UBLKCP = cp.async.bulk (1-D bulk copies), SYNCS = mbarrier, FFMA2 = packed fma.rn.f32x2, LD.E/ST.E = GENERIC loads/stores
(should only appear where a pointer really is generic: peer mailboxes, staging reads outside the hot loops)."""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "soft-labeled-contrastive-learning_b200", "slcl", "libslcl.so")
PAT = [("UTCHMMA", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"),
       ("UBLKCP", r"\bUBLKCP"), ("UTCBAR", r"\bUTCBAR"), ("SYNCS", r"\bSYNCS"), ("FFMA2", r"\bFFMA2"), ("FFMA", r"\bFFMA\b"),
       ("MUFU.EX2", r"\bMUFU\.EX2"), ("LDS", r"\bLDS"), ("STS", r"\bSTS"), ("LDG", r"\bLDG"), ("STG", r"\bSTG"),
       ("LDGSTS", r"\bLDGSTS"), ("LD.E", r"\bLD\.E"), ("ST.E", r"\bST\.E")]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            order.append(cur)
            continue
        if cur is None or "/*" not in line:
            continue
        for name, pat in PAT:
            if re.search(pat, line):
                counts[cur][name] += 1
    dm = demangle(order)
    digest = hashlib.sha256(open(LIB, "rb").read()).hexdigest()[:16]
    print(f"# cuobjdump -sass summary of slcl/libslcl.so (sha256 {digest}, {len(order)} kernels); columns = instruction counts")
    print("# " + " ".join(f"{n:>8s}" for n, _ in PAT) + "  kernel")
    total = collections.Counter()
    for fn in order:
        c = counts[fn]
        total.update(c)
        name = re.sub(r"slcl::\(anonymous namespace\)::", "", dm.get(fn, fn))
        name = re.sub(r"\(CUtensorMap_st.*", "(...)", name)
        name = re.sub(r"\((anonymous namespace::)?[A-Za-z]+Args.*", "(...)", name)
        print("  " + " ".join(f"{c.get(n, 0):8d}" for n, _ in PAT) + "  " + name[:110])
    print("# " + " ".join(f"{total.get(n, 0):8d}" for n, _ in PAT) + "  TOTAL")


if __name__ == "__main__":
    sys.exit(main())

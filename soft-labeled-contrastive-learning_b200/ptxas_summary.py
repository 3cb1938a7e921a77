"""Summarise `-Xptxas -v` logs from build/: registers, spills, smem per kernel."""
import glob, os, re, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
rows = []
for log in sorted(glob.glob(os.path.join(here, "build", "libslcl", "*.ptxas.log"))):
    txt = open(log).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?Function properties for \S+\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?", txt, re.S):
        rows.append((os.path.basename(log)[:-10], m.group(1), int(m.group(5)), int(m.group(2)), int(m.group(3)), int(m.group(6) or 0)))
names = subprocess.run(["c++filt"] + [r[1] for r in rows], capture_output=True, text=True).stdout.splitlines()
for r, n in zip(rows, names):
    n = re.sub(r"\(.*", "", n).replace("slcl::(anonymous namespace)::", "").replace("void ", "")
    print(f"{r[0]:14s} regs={r[2]:3d} stack={r[3]:4d} spill_st={r[4]:4d} smem={r[5]:6d}  {n}")

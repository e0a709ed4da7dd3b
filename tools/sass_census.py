"""Census of the Blackwell-specific SASS in the built library (runs without a GPU): per kernel, how many TMA bulk copies
(UBLKCP), mbarrier operations (SYNCS), elected-thread issues (ELECT), warp-wide reductions (CREDUX: what
__reduce_min_sync / __reduce_add_sync compile to on sm_100a), shuffles, match / vote, asynchronous copies (LDGSTS) and
atomics / reductions to global and shared memory it contains, plus its register count.

    python tools/sass_census.py > profiles/r2_sass_census.txt
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pylidar_slam_b200", "libplslam_b200.so")
OPS = ("UBLKCP", "SYNCS", "ELECT", "CREDUX", "REDUX", "SHFL", "MATCH", "VOTE", "VOTEU", "LDGSTS", "ATOMG", "ATOMS", "REDG",
       "UTMALDG", "UTCMMA", "MEMBAR", "FENCE")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    clean = []
    for n in out:
        n = n.replace("(anonymous namespace)::", "").replace("pls::", "")
        n = re.sub(r"\(.*", "", n)
        n = re.sub(r"^void ", "", n)
        clean.append(n)
    return clean


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+)", line)
        if m and cur:
            regs[cur] = int(m.group(1))
    counts, order, fn = {}, [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            order.append(fn)
            counts[fn] = {}
            continue
        if fn is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for o in OPS:
                if op == o:
                    counts[fn][o] = counts[fn].get(o, 0) + 1
    names = demangle(order)
    arch = set(re.findall(r"arch = (sm_\w+)", sass))
    print(f"# SASS census of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass; architectures present: {', '.join(sorted(arch))})")
    print(f"# {len(order)} kernels; columns: registers, then instruction counts of {', '.join(OPS)} (zeros omitted)")
    for fn, name in sorted(zip(order, names), key=lambda t: t[1]):
        ops = " ".join(f"{o}={c}" for o, c in counts[fn].items())
        print(f"{name:70s} regs={regs.get(fn, '?'):>3}  {ops}")
    tot = {o: sum(c.get(o, 0) for c in counts.values()) for o in OPS}
    print("# totals: " + " ".join(f"{o}={n}" for o, n in tot.items() if n))


if __name__ == "__main__":
    sys.exit(main())

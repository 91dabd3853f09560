"""Which kernels of a new build differ from a reference build (e.g. the last hardware-validated one)?  Compares the SASS of
every function of two libhold_b200.so files, addresses normalised, template-renamed tcgen05 kernels matched.
usage: python tools/sass_diff.py /path/to/validated.so [hold_b200/libhold_b200.so]"""
import re
import subprocess
import sys


def funcs(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    d, cur, buf = {}, None, []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if cur:
                d[cur] = buf
            cur, buf = m.group(1), []
        elif cur:
            mm = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
            if mm:
                buf.append(re.sub(r"0x[0-9a-f]+", "A", mm.group(1)))
    if cur:
        d[cur] = buf
    return d


def main():
    ref, new = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "hold_b200/libhold_b200.so")
    a, b = funcs(ref), funcs(new)
    dem = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()[:100]
    same = diff = 0
    for k in sorted(a):
        kb = k if k in b else None
        if kb is None:
            m = re.match(r"(_ZN4hold8k_mlp_tcILi\d)E(.*)", k)       # k_mlp_tc<M> -> k_mlp_tc<M, false>
            kb = (m.group(1) + "ELb0E" + m.group(2)) if m and (m.group(1) + "ELb0E" + m.group(2)) in b else None
        if kb is None:
            print("GONE ", dem(k))
        elif a[k] == b[kb]:
            same += 1
        else:
            diff += 1
            print(f"DIFF  {dem(k)}  ({len(a[k])} -> {len(b[kb])} instructions)")
    for k in sorted(set(b) - set(a)):
        if not re.match(r"_ZN4hold8k_mlp_tcILi\dELb0E", k):
            print("NEW  ", dem(k))
    print(f"{same} kernels identical, {diff} differ")


if __name__ == "__main__":
    main()

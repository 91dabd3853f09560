"""Which kernels of a new build differ from a reference build (e.g. the last hardware-validated one)?  Compares the SASS of
every function of two libhold_b200.so files, addresses normalised, template-renamed tcgen05 kernels matched.
usage: python tools/sass_diff.py /path/to/validated.so [hold_b200/libhold_b200.so]
       python tools/sass_diff.py --hash lib.so > profiles/rNN_validated_sass_hashes.txt
       python tools/sass_diff.py --check profiles/rNN_validated_sass_hashes.txt [lib.so]"""
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


def hashes(lib):
    import hashlib

    return {k: hashlib.sha1("\n".join(v).encode()).hexdigest() for k, v in funcs(lib).items()}


def main():
    if sys.argv[1] == "--hash":          # python tools/sass_diff.py --hash lib.so > file : one "sha1 mangled-name" line per kernel
        for k, h in sorted(hashes(sys.argv[2]).items()):
            print(h, k)
        return
    if sys.argv[1] == "--check":         # python tools/sass_diff.py --check hashes.txt [lib.so] : kernels whose SASS changed
        want = dict(reversed(l.split()) for l in open(sys.argv[2]) if l.strip())
        have = hashes(sys.argv[3] if len(sys.argv) > 3 else "hold_b200/libhold_b200.so")
        same = 0
        for k, h in sorted(want.items()):
            cand = [k] + [re.sub(r"(k_mlp_tcILi\d)E", r"\1" + mid, k) for mid in ("ELb0E", "ELb0ELb0E")]
            got = next((have[c] for c in cand if c in have), None)
            if got == h:
                same += 1
            else:
                print("CHANGED" if got else "GONE   ", k)
        print(f"{same} of {len(want)} kernels identical to the recorded build")
        return
    ref, new = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "hold_b200/libhold_b200.so")
    a, b = funcs(ref), funcs(new)
    dem = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()[:100]
    same = diff = 0
    for k in sorted(a):
        kb = k if k in b else None
        if kb is None:
            m = re.match(r"(_ZN4hold8k_mlp_tcILi\d)E(.*)", k)       # k_mlp_tc<M> -> k_mlp_tc<M, false[, false]>
            for mid in ("ELb0E", "ELb0ELb0E"):
                if m and (m.group(1) + mid + m.group(2)) in b:
                    kb = m.group(1) + mid + m.group(2)
        if kb is None:
            print("GONE ", dem(k))
        elif a[k] == b[kb]:
            same += 1
        else:
            diff += 1
            print(f"DIFF  {dem(k)}  ({len(a[k])} -> {len(b[kb])} instructions)")
    for k in sorted(set(b) - set(a)):
        if not re.match(r"_ZN4hold8k_mlp_tcILi\d(ELb0E|ELb0ELb0E)E", k):
            print("NEW  ", dem(k))
    print(f"{same} kernels identical, {diff} differ")


if __name__ == "__main__":
    main()

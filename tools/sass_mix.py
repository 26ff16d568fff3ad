"""Instruction mix of the innermost loop (or whole body) of the kernels in a cubin/executable whose mangled name contains
a pattern:  python tools/sass_mix.py <binary> <name-substring> [--whole]"""
import collections, re, subprocess, sys

def main():
    binary, pat = sys.argv[1], sys.argv[2]
    whole = "--whole" in sys.argv
    txt = subprocess.run(["cuobjdump", "-sass", binary], capture_output=True, text=True).stdout
    for f in re.split(r"\n\s+Function : ", txt)[1:]:
        name = f.split("\n")[0]
        if pat not in name:
            continue
        lines = [l for l in f.split("\n") if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l)]
        addr = lambda l: int(re.match(r"\s+/\*([0-9a-f]{4,5})\*/", l).group(1), 16)
        body = lines
        if not whole:
            best = None
            for l in lines:
                m = re.search(r"BRA.*0x([0-9a-f]+)", l)
                if m and int(m.group(1), 16) < addr(l):
                    lo, hi = int(m.group(1), 16), addr(l)
                    if best is None or hi - lo > best[1] - best[0]:
                        best = (lo, hi)
            if best:
                body = [x for x in lines if best[0] <= addr(x) <= best[1]]
        def op(x):
            t = x.split()
            o = t[2] if t[1].startswith("@") else t[1]
            return o.rstrip(";")
        full = collections.Counter(op(x) for x in body)
        short = collections.Counter(re.sub(r"\..*", "", k) for k in full.elements())
        print(name, "instructions:", len(body))
        print("  ", short.most_common())
        print("  ", [(k, v) for k, v in full.most_common() if k.startswith(("IMAD", "DFMA", "DADD", "DMUL"))])

main()

"""Static SASS opcode census of the built library (no GPU needed): `cuobjdump -sass` per kernel,
counts of the tensor / TMA / TMEM opcodes that prove the tcgen05 path plus the common memory and
math opcodes.  usage: sass_census.py [out.md]"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "atomai_b200", "libatomai_b200.so")
TENSOR = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "SYNCS", "USETMAXREG"]
OTHER = ["UTCATOMSWS", "LDG", "STG", "LDS", "STS", "FFMA", "MUFU", "SHFL"]


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    res = []
    for n in out:
        n = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", n)
        n = re.sub(r"\(.*$", "", n)
        res.append(n.replace("(int)", ""))
    return res


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_n"] += 1
            cur[op.split(".")[0]] += 1
    names = demangle(list(kernels))
    rows = sorted(zip(names, kernels.values()), key=lambda t: -t[1]["_n"])
    lines = ["# round 2 — SASS census of atomai_b200/libatomai_b200.so (sm_100a)", "",
             "`python tools/sass_census.py` (`cuobjdump -sass`, static instruction counts per kernel). "
             "`UTCHMMA` = tcgen05.mma (TF32 `kind::tf32` and the bf16 `kind::f16` correction MMA of tf32x3), "
             "`LDTM` = tcgen05.ld (TMEM -> registers), `UTMALDG` = cp.async.bulk.tensor (TMA tile load), "
             "`UBLKCP` = cp.async.bulk (weight blobs), `UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier ops, "
             "`USETMAXREG` = setmaxnreg.", "",
             "| kernel | SASS instr | tensor / TMA / TMEM opcodes | other |", "|---|---:|---|---|"]
    for n, c in rows:
        if c["_n"] < 300:
            continue
        t = ", ".join(f"{k} {c[k]}" for k in TENSOR if c[k]) or "-"
        o = ", ".join(f"{k} {c[k]}" for k in OTHER if c[k])
        lines.append(f"| `{n}` | {c['_n']} | {t} | {o} |")
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text)
    print(text[:3000])


if __name__ == "__main__":
    main()

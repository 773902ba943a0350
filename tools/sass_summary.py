"""What the shipped device code contains, read from the built library without a GPU:

    python tools/sass_summary.py > profiles/r02_sass.md

* `cuobjdump -sass neuralmonkey_b200/libnmb200.so`: per kernel family (demangled name without template arguments)
  the number of instances and the counts of the SASS mnemonics that prove a Blackwell-native path
  (/opt/skills/guides/B200_PROFILING.md, "What proves a Blackwell-native kernel"): `UTC*MMA` = tcgen05.mma,
  `LDTM` / `STTM` = tcgen05.ld / st, `UTMALDG` / `UTMASTG` = TMA tensor copies, `UBLKCP` = cp.async.bulk (1-D TMA),
  `HMMA` = the legacy mma.sync path (none expected), plus cluster / DSMEM evidence (`UCGABAR*` cluster barriers,
  `ST*.*.CLUSTER`-space stores show as `STS`/`STAS`; counted: `UCGABAR`, `STAS`, `SYNCS`, `MUFU`).
* `nvcc -Xptxas -v` over every csrc/*.cu: registers per thread, spill bytes, static shared memory per kernel;
  kernels with spills are listed individually."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "neuralmonkey_b200", "libnmb200.so")
CSRC = os.path.join(ROOT, "neuralmonkey_b200", "csrc")
MNEMONICS = [("UTC*MMA", r"\bUTC\w*MMA\b"), ("LDTM", r"\bLDTM\b"), ("STTM", r"\bSTTM\b"),
             ("UTMALDG", r"\bUTMALDG\b"), ("UTMASTG", r"\bUTMASTG\b"), ("UBLKCP", r"\bUBLKCP\b"),
             ("SYNCS", r"\bSYNCS\b"), ("UCGABAR", r"\bUCGABAR\w*"), ("STAS", r"\bSTAS\b"), ("MUFU", r"\bMUFU\b"),
             ("HMMA", r"\bHMMA\b")]


def demangle(names):
    res = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True)
    return res.stdout.splitlines()


def family(demangled):
    name = re.sub(r"^void ", "", demangled)
    name = name.split("(")[0]
    return re.sub(r"<.*", "", name)


def sass_table():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, name = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = collections.Counter()
            continue
        if name is None or not re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):      # instruction lines carry their address
            continue
        for label, pattern in MNEMONICS:
            if re.search(pattern, line):
                kernels[name][label] += 1
        kernels[name]["instructions"] += 1
    mangled = list(kernels)
    families = collections.OrderedDict()
    for m_name, d_name in zip(mangled, demangle(mangled)):
        fam = families.setdefault(family(d_name), {"instances": 0, "counts": collections.Counter()})
        fam["instances"] += 1
        fam["counts"].update(kernels[m_name])
    return families, len(mangled)


def ptxas_table():
    rows, spills = [], []
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xptxas", "-v",
             "-c", "-o", "/dev/null"]
    for src in sorted(f for f in os.listdir(CSRC) if f.endswith(".cu")):
        res = subprocess.run(["nvcc"] + flags + [os.path.join(CSRC, src)], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(res.stderr)
        text = res.stderr
        entries = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\n(?:.*\n)*?.*?(\d+) bytes stack frame, "
                             r"(\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers"
                             r"(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?", text)
        regs = [int(e[4]) for e in entries]
        rows.append((src, len(entries), max(regs) if regs else 0, sum(1 for e in entries if int(e[2]) or int(e[3]))))
        names = demangle([e[0] for e in entries]) if entries else []
        for e, d_name in zip(entries, names):
            if int(e[2]) or int(e[3]):
                spills.append((src, d_name[:110], int(e[4]), int(e[1]), int(e[2]), int(e[3])))
    return rows, spills


def main():
    families, total = sass_table()
    print("# Device code of `libnmb200.so` (sm_100a): SASS evidence and ptxas resource usage\n")
    print("Generated here, without a GPU, by `python tools/sass_summary.py` from the in-tree library the GPU box loads "
          "(`cuobjdump -sass`) and from `nvcc -Xptxas -v` over `csrc/*.cu` (CUDA 12.9).  {} kernels in {} "
          "families.  `UTC*MMA` = `tcgen05.mma`, `LDTM` / `STTM` = `tcgen05.ld` / `tcgen05.st`, `UTMALDG` / `UTMASTG` = "
          "TMA tensor copies (`cp.async.bulk.tensor`), `UBLKCP` = 1-D `cp.async.bulk`, `SYNCS` = mbarrier / "
          "transaction-barrier operations, `UCGABAR` = cluster barriers, `STAS` = `st.async` into a peer CTA's shared "
          "memory, `HMMA` = the legacy `mma.sync` path (none).\n".format(total, len(families)))
    labels = [l for l, _ in MNEMONICS]
    print("| kernel family | instances | instructions | " + " | ".join(labels) + " |")
    print("|---|---|---|" + "---|" * len(labels))
    order = sorted(families.items(), key=lambda kv: (-kv[1]["counts"]["UTC*MMA"], -kv[1]["counts"]["UBLKCP"],
                                                     -kv[1]["counts"]["SYNCS"], kv[0]))
    for name, fam in order:
        c = fam["counts"]
        print("| `{}` | {} | {} | ".format(name, fam["instances"], c["instructions"])
              + " | ".join(str(c[l]) if c[l] else "" for l in labels) + " |")
    totals = collections.Counter()
    for fam in families.values():
        totals.update(fam["counts"])
    print("| **all** | {} | {} | ".format(total, totals["instructions"])
          + " | ".join(str(totals[l]) for l in labels) + " |")
    rows, spills = ptxas_table()
    print("\n## ptxas -v\n\n| source | kernels | max registers / thread | kernels with spills |\n|---|---|---|---|")
    for src, n, regs, spilled in rows:
        print("| `csrc/{}` | {} | {} | {} |".format(src, n, regs, spilled))
    if spills:
        print("\nKernels with register spills:\n\n| source | kernel | registers | stack bytes | spill stores | spill loads |"
              "\n|---|---|---|---|---|---|")
        for row in spills:
            print("| `{}` | `{}` | {} | {} | {} | {} |".format(*row))
    else:
        print("\nNo kernel spills registers.")


if __name__ == "__main__":
    sys.exit(main())

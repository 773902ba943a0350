"""Check that a refactoring left the device code of existing kernels untouched.

    cuobjdump -sass old/gemm_tc.o > old.sass      # before the change
    cuobjdump -sass new/gemm_tc.o > new.sass      # after
    python tools/sass_identity.py old.sass new.sass

Kernels are matched by their demangled name with trailing default template arguments ignored (a new
defaulted template parameter changes the mangled name, not the code); the instruction text AND the
encodings of every matched kernel must be identical.  Used when `tc_gemm_kernel` got its element-size
parameter and the trailing `TcExt` argument (20 of 20 TF32 instances identical) and when the tiled MHA
kernels got their dropout variants (10 of 10), so the GPU-verified paths did not change although no
GPU was available to re-run their tests."""
import re
import subprocess
import sys


def split(path):
    funcs, name, buf = {}, None, []
    for line in open(path):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                funcs[name] = buf
            name, buf = m.group(1), []
        elif name is not None and "/*" in line:
            buf.append(" ".join(line.split()))        # cuobjdump aligns its columns per file
    if name:
        funcs[name] = buf
    return funcs


def key(mangled):
    demangled = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    m = re.match(r"(?:void )?([\w:]+)<(.*?)>\(", demangled)
    if m:
        return m.group(1), tuple(a.strip() for a in m.group(2).split(","))
    return re.sub(r"^void ", "", demangled).split("(")[0], ()       # not a template (yet)


def main():
    old, new = split(sys.argv[1]), split(sys.argv[2])
    new_keys = {key(k): v for k, v in new.items()}
    same = 0
    for mangled, body in old.items():
        name, args = key(mangled)
        match = [v for (n, a), v in new_keys.items() if n == name and a[:len(args)] == args and
                 all(x in ("4", "false") for x in a[len(args):])]   # trailing args = the new defaults
        if len(match) == 1 and match[0] == body:
            same += 1
        else:
            print("DIFFERENT or missing:", name, args)
    print("identical: {} of {}".format(same, len(old)))
    return 0 if same == len(old) else 1


if __name__ == "__main__":
    sys.exit(main())

"""Diagnostic: probe MN-major tf32 smem descriptor / TMA swizzle combinations (one process each)."""
import itertools, os, subprocess, sys
CODE = r'''
import sys, torch
sys.path.insert(0, ".")
from neuralmonkey_b200 import lib, ops
torch.manual_seed(0)
res = []
for (ta, tb) in ((True, True), (False, False), (True, False)):
    m, n, k = 256, 256, 96
    a = torch.randn(k, m) if ta else torch.randn(m, k)
    b = torch.randn(n, k) if tb else torch.randn(k, n)
    out = torch.zeros(m, n, device="cuda")
    ops.gemm(a.cuda(), b.cuda(), out, trans_a=ta, trans_b=tb, backend=lib.GEMM_TC)
    ref = (a.t() if ta else a).double() @ (b.t() if tb else b).double()
    err = float((out.double().cpu() - ref).norm() / ref.norm())
    res.append("%.2e" % err)
print(" ".join(res))
'''
combos = []
for layout, swz in ((1, 4), (2, 3), (1, 3), (2, 4), (1, 6), (1, 5)):
    for sbo in (512, 1024):
        for lbo in (4096,):
            for kadv in (1024, 32, 512):
                combos.append((layout, swz, sbo, lbo, kadv))
for layout, swz, sbo, lbo, kadv in combos:
    env = dict(os.environ, NMB200_MN_LAYOUT=str(layout), NMB200_MN_SWIZZLE=str(swz), NMB200_MN_SBO=str(sbo),
               NMB200_MN_LBO=str(lbo), NMB200_MN_KADV=str(kadv))
    try:
        r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True, timeout=120)
        out = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ("ERR " + r.stderr.strip()[-200:])
    except subprocess.TimeoutExpired:
        out = "TIMEOUT"
    print("layout=%d swizzle=%d sbo=%d lbo=%d kadv=%d -> [TT FF TF] %s" % (layout, swz, sbo, lbo, kadv, out), flush=True)

"""GPU probe: tensor-core GRU vs exact engine (values + timing at the bench shape)."""
import sys, time
import torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from neuralmonkey_b200 import lib, ops

def run(mode, B, T, E, H, reverse=False, lengths=None, seed=0, bwd=True):
    lib.call("nm_gru_set_mode", mode)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, E, generator=g).cuda().requires_grad_(True)
    wg = (torch.randn(E + H, 2 * H, generator=g) * 0.1).cuda().requires_grad_(True)
    bg = torch.ones(2 * H).cuda().requires_grad_(True)
    wc = (torch.randn(E + H, H, generator=g) * 0.1).cuda().requires_grad_(True)
    bc = torch.zeros(H).cuda().requires_grad_(True)
    h0 = (torch.randn(B, H, generator=g) * 0.5).cuda().requires_grad_(True)
    ld = lengths.to(torch.int32).cuda() if lengths is not None else None
    st, fin, raw = ops.gru_layer(x, wg, bg, wc, bc, h0, lengths=ld, reverse=reverse)
    grads = None
    if bwd:
        ds = torch.randn(B, T, H, generator=g).cuda()
        (st * ds).sum().backward()
        grads = [t.grad.clone() for t in (x, wg, wc, h0)]
    torch.cuda.synchronize()
    return st.detach(), fin.detach(), grads

for (B, T, E, H) in [(5, 6, 11, 7), (9, 4, 32, 32), (70, 3, 16, 64), (256, 50, 300, 300), (3, 5, 8, 100)]:
    for reverse in (False, True):
        lengths = torch.randint(1, T + 1, (B,), generator=torch.Generator().manual_seed(1))
        lengths[0] = T
        a = run(1, B, T, E, H, reverse, lengths)
        b = run(0, B, T, E, H, reverse, lengths)
        e1 = float((a[0] - b[0]).abs().max()); e2 = float((a[1] - b[1]).abs().max())
        ge = [float((x - y).norm() / (x.norm() + 1e-30)) for x, y in zip(a[2], b[2])]
        print("B%d T%d E%d H%d rev=%d  states %.2e final %.2e grads %s" % (B, T, E, H, reverse, e1, e2, ["%.1e" % v for v in ge]), flush=True)

# timing of the raw ABI call at the bench shape
B, T, H = 256, 50, 300
xproj = torch.randn(B, T, 3 * H).cuda() * 0.3
wgh = (torch.randn(H, 2 * H) * 0.1).cuda(); wch = (torch.randn(H, H) * 0.1).cuda()
outs = [torch.empty(B, T, H).cuda() for _ in range(2)]
fin = torch.empty(B, H).cuda(); gates = torch.empty(B, T, 3 * H).cuda()
hprev = torch.empty(B, T, H).cuda(); rh = torch.empty(B, T, H).cuda()
for mode in (1, 0):
    lib.call("nm_gru_set_mode", mode)
    def f():
        lib.call("nm_gru_seq_fwd", lib.ptr(xproj), lib.ptr(wgh), lib.ptr(wch), None, None, None, 0, lib.ptr(outs[0]),
                 lib.ptr(outs[1]), lib.ptr(fin), lib.ptr(gates), lib.ptr(hprev), lib.ptr(rh), B, T, H, 0, lib.stream())
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    print("mode %d fwd %.1f us" % (mode, e0.elapsed_time(e1) * 100), flush=True)
    dst = torch.randn(B, T, H).cuda() * 1e-3; dxp = torch.empty(B, T, 3 * H).cuda(); work = torch.empty(2 * B * H).cuda()
    def gb():
        lib.call("nm_gru_seq_bwd", lib.ptr(wgh), lib.ptr(wch), None, None, 0, lib.ptr(gates), lib.ptr(hprev), lib.ptr(dst),
                 None, None, lib.ptr(dxp), None, lib.ptr(work), B, T, H, 0, lib.stream())
    for _ in range(3): gb()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10): gb()
    e1.record(); torch.cuda.synchronize()
    print("mode %d bwd %.1f us" % (mode, e0.elapsed_time(e1) * 100), flush=True)

# cycle profile of the tensor-core forward (CTA 0, thread 0)
prof = torch.zeros(8, dtype=torch.int64).cuda()
lib.call("nm_gru_debug_profile", lib.ptr(prof))
lib.call("nm_gru_set_mode", 0)
f(); torch.cuda.synchronize()
lib.call("nm_gru_debug_profile", None)
names = ["wait+mma+dump1", "items1", "sync1", "send1", "wait+mma+dump2", "items2", "sync2", "send2"]
tot = float(prof.sum())
print("cycles/step:", {n: int(v) // T for n, v in zip(names, prof.tolist())}, "total/step", int(tot) // T)

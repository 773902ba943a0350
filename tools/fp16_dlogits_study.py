"""CPU study for DESIGN.md section 8 item 1: does storing the vocabulary gradient (dlogits) in fp16 keep the
accuracy of today's TF32 path?  Emulates both on a scaled-down logits problem and prints the relative
error of dW and dX against fp64.  Result (2048 x 300 x 8192): TF32 path 3.8e-4; fp16 of the UNNORMALISED
p - onehot (the 1/count applied in the consumer's epilogue) 2.5e-4; fp16 of the normalised gradient 2e-2
(99.8 % of the entries underflow); bf16 1.7e-3."""
import torch, math
torch.manual_seed(0)
M,K,V = 2048, 300, 8192   # rows (tokens), O, vocab (scaled down)
X = torch.randn(M,K)*0.5
W = torch.randn(K,V)*0.05
b = torch.zeros(V)
t = torch.randint(0,V,(M,))
count = 12800.0
def tf32(x):
    # round to 10-bit mantissa (tf32), keep exponent range
    xi = x.view(torch.int32)
    xi = (xi + 0x1000) & ~0x1FFF
    return xi.view(torch.float32)
logits64 = X.double()@W.double()
p64 = torch.softmax(logits64,-1)
d64 = p64.clone(); d64[torch.arange(M),t] -= 1; d64 /= count
dW64 = X.double().t()@d64; dX64 = d64@W.double().t()
# current path: logits via tf32 operands, dlogits fp32, consumers with tf32 operands
logits = tf32(X)@tf32(W)
p = torch.softmax(logits,-1); d = p.clone(); d[torch.arange(M),t]-=1; d/=count
dW_tf32 = (tf32(X).t().double()@tf32(d).double()); dX_tf32 = tf32(d).double()@tf32(W).double().t()
def rel(a,b): return float((a-b).norm()/b.norm())
print("tf32 path   dW rel %.2e  dX rel %.2e"%(rel(dW_tf32,dW64), rel(dX_tf32,dX64)))
# fp16 dlogits stored normalised (tiny values)
for name, scale in (("fp16 normalised", 1.0), ("fp16 x count (p - onehot)", count), ("fp16 x 1024", 1024.0)):
    dh = (d*scale).half()
    dd = dh.double()/scale
    dW = X.half().double().t()@dd; dX = dd@W.half().double().t()
    print("%-28s dW rel %.2e  dX rel %.2e  zeros %.1f%%  min|nz| %.1e"%(name, rel(dW,dW64), rel(dX,dX64), 100*float((dh==0).float().mean()), float(dh[dh!=0].abs().min())))
# bf16 for comparison
dd=(d*count).bfloat16().double()/count
print("bf16 x count                 dW rel %.2e  dX rel %.2e"%(rel(X.bfloat16().double().t()@dd,dW64), rel(dd@W.bfloat16().double().t(),dX64)))

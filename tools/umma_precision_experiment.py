import ctypes as C, torch, sys
sys.path.insert(0,'/root/repo')
from elg_b200 import _lib
def run(a, b, terms, alias=0):
    d = torch.zeros(128, b.shape[0], device='cuda')
    _lib.check(_lib.lib.elg_selftest_umma(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(d.data_ptr()), a.shape[0], b.shape[0], a.shape[1], alias, terms, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize(); return d[:a.shape[0]].double()
g = torch.Generator().manual_seed(1)
a = torch.randn(128, 128, generator=g).cuda(); b = (torch.randn(128, 128, generator=g)*0.1).cuda()
ref = a.double() @ b.double().T
scale = ref.abs().mean()
def rep(name, d): print('%-28s max %.2e  mean %.2e (relative to mean |ref|)' % (name, (d-ref).abs().max()/scale, (d-ref).abs().mean()/scale))
rep('fp32 torch matmul (no tf32)', (a @ b.T).double())
for t in (1,2,3,4): rep('umma terms=%d' % t, run(a, b, t))
# exactly representable inputs (11 significant bits): isolates accumulation error
a2 = a.half().float(); b2 = b.half().float(); ref = a2.double() @ b2.double().T
rep('fp16-exact inputs: fp32 matmul', (a2 @ b2.T).double()); rep('fp16-exact inputs: umma t=1', run(a2, b2, 1))
# quantization only: hi+lo representation error effect (double precision product of reconstructed operands)
ah = a.half().float(); al = (a - ah).half().float(); bh = b.half().float(); bl = (b - bh).half().float()
ref = a.double() @ b.double().T
rep('exact product of (hi+lo) operands', (ah.double()+al.double()) @ (bh.double()+bl.double()).T)
rep('3-term exact', ah.double()@bh.double().T + ah.double()@bl.double().T + al.double()@bh.double().T)
print('--- power-of-two pre-scaling (keeps the lo halves out of the fp16 subnormal range) ---')
ref = a.double() @ b.double().T
for sa, sb in ((0, 0), (4, 4), (7, 10), (8, 12)):
    d = run(a * 2.0**sa, b * 2.0**sb, 3) * 2.0**-(sa+sb)
    rep('umma 3 terms, scale 2^%d x 2^%d' % (sa, sb), d)
print('--- separate accumulator for the cross terms (terms=5) ---')
rep('umma hi*hi | cross terms separate', run(a, b, 5))
a3 = torch.randn(100, 128, generator=g).cuda() * 3; b3 = torch.randn(112, 128, generator=g).cuda() * 0.3
ref = a3.double() @ b3.double().T; scale = ref.abs().mean()
rep('other data: fp32 matmul', (a3 @ b3.T).double()); rep('other data: umma 3 terms', run(a3, b3, 3)); rep('other data: umma split acc', run(a3, b3, 5))
print('--- A operand in tensor memory (tcgen05.mma TS form, terms=6) ---')
rep('other data: umma TS split acc', run(a3, b3, 6))
ref = a.double() @ b.double().T; scale = ref.abs().mean()
rep('first data: umma TS split acc', run(a, b, 6))
print('--- single accumulator, cross terms issued first, hi*hi last (terms=7) ---')
rep('first data: cross-first single acc', run(a, b, 7))
ref = a3.double() @ b3.double().T; scale = ref.abs().mean()
rep('other data: cross-first single acc', run(a3, b3, 7))

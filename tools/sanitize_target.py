"""Small-shape run of the kernels that alias TMEM columns / use TMA reduce-add, for compute-sanitizer (tools/sanitize.sh).
what = attn | gemm | model"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from osu_dreamer_b200 import lib  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else 'attn'
g = torch.Generator().manual_seed(0)
if what == 'attn':
    for B, L in ((1, 320), (2, 128)):  # ragged (320 = 2.5 q tiles) and multi-batch
        qkv = torch.randn(B * L, 3072, generator=g).cuda().to(torch.bfloat16)
        dy = torch.randn(B * L, 1024, generator=g).cuda().to(torch.bfloat16)
        bound = torch.tensor([14.0], device='cuda')
        y, lse = lib.attn_fwd(qkv, B, L, bound_log2=bound, variant=7)
        y2, lse2 = lib.attn_fwd(qkv, B, L, variant=7)  # online-softmax path of the same kernel
        dqkv = lib.attn_bwd_fused(qkv, y, dy, lse, B, L)
        dqkv2 = lib.attn_bwd(qkv, y, dy, lse, B, L)
        torch.cuda.synchronize()
        print('attn', B, L, float(y.float().abs().max()), float(dqkv.float().abs().max()), float((dqkv.float() - dqkv2.float()).abs().max()))
elif what == 'gemm':
    T = 320
    x = torch.randn(T, 512, generator=g).cuda().to(torch.bfloat16)
    w = torch.randn(3072, 512, generator=g).cuda().to(torch.bfloat16)
    c = torch.empty(T, 3072, device='cuda', dtype=torch.bfloat16)
    lib.gemm(x, w, c)
    cf = torch.zeros(512, 3072, device='cuda')
    lib.gemm(x, c, cf, a_major=lib.MAJOR_MN, b_major=lib.MAJOR_MN, epi=lib.EPI_ATOMIC, split_k=2)  # wgrad shape, split-K reduce-add
    rope = lib.rope_table(160, 'cuda')
    qn = torch.ones(64, device='cuda')
    out = lib.qkv_proj(x, w, torch.zeros(3072, device='cuda'), qn, qn, rope, 160)
    torch.cuda.synchronize()
    print('gemm', float(c.float().abs().max()), float(cf.abs().max()), float(out.float().abs().max()))
else:
    from oracle import denoiser_oracle as O
    from osu_dreamer_b200.denoiser import DiffusionModel, default_args
    m = DiffusionModel(6, 128, 32, default_args())
    m.load_state_dict(O.make_state_dict(1234))
    m = m.cuda().train()
    inp = O.make_inputs(1, 192, seed=3)
    u, v = m(inp['h'].cuda(), inp['s'].cuda(), inp['x0'].cuda())
    (v.square().mean() + u.square().mean()).backward()
    m.precision = 'fp32'
    with torch.no_grad():
        m.eval()(inp['h'].cuda(), inp['s'].cuda(), inp['x0'].cuda())
    torch.cuda.synchronize()
    print('model', float(v.abs().max()))

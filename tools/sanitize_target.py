"""Small-shape run of the kernels that alias TMEM columns / use TMA reduce-add, for compute-sanitizer (tools/sanitize.sh).
what = attn | gemm | model | new (the kernels added late in round 2: CTA-pair GEMM, forward variants 17 / 18, the compact-grid
backward fallback, the tensor-core latent block)"""
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
elif what == 'new':
    # the kernels added late in round 2; run with OSD_GEMM_PAIR=1 so that the small shapes below take the CTA-pair GEMM
    M = 256 * 17 + 40  # 18 pair tiles x 4 column tiles >= 0.9 x 74 pairs; the last tile is ragged
    a = torch.randn(M, 512, generator=g).cuda().to(torch.bfloat16)
    w = torch.randn(1024, 512, generator=g).cuda().to(torch.bfloat16)
    bias = torch.randn(1024, generator=g).cuda()
    c = lib.gemm(a, w, torch.empty(M, 1024, device='cuda', dtype=torch.bfloat16), bias=bias)
    cf = lib.gemm(a, w.T.contiguous(), torch.empty(M, 1024, device='cuda'), b_major=lib.MAJOR_MN)
    dw = torch.zeros(1024, 512, device='cuda')
    lib.gemm(c, a, dw, a_major=lib.MAJOR_MN, b_major=lib.MAJOR_MN, epi=lib.EPI_ATOMIC, split_k=lib.gemm_split_k(1024, 512, M))
    rope = lib.rope_table(M // 8, 'cuda')
    qn = torch.ones(64, device='cuda')
    w3 = torch.randn(3072, 512, generator=g).cuda().to(torch.bfloat16)
    out = lib.qkv_proj(a[:M // 8 * 8], w3, torch.zeros(3072, device='cuda'), qn, qn, rope, M // 8)
    torch.cuda.synchronize()
    print('pair gemm', float((c.float() - (a.float() @ w.float().T + bias)).abs().max()), float(cf.abs().max()), float(dw.abs().max()),
          float(out.float().abs().max()))
    for B, L in ((1, 600), (2, 256)):  # forward variants 18 (two q tiles, sixteen softmax warps) and 17 (multicast cluster)
        qkv = torch.randn(B * L, 3072, generator=g).cuda().to(torch.bfloat16)
        bound = torch.tensor([14.0], device='cuda')
        for variant in (18, 17):
            y, lse = lib.attn_fwd(qkv, B, L, bound_log2=bound, variant=variant)
            y2, _ = lib.attn_fwd(qkv, B, L, variant=variant)  # online-softmax path
            torch.cuda.synchronize()
            print('attn fwd', variant, B, L, float((y.float() - y2.float()).abs().max()))
    B, L = 3, 1920  # single-pass backward whose gate opens: the two-kernel fallback on its compact grid (CTAs walk 2-3 items)
    qkv = torch.randn(B * L, 3072, generator=g)
    qkv[3, :1024] *= 40.0
    qkv = qkv.cuda().to(torch.bfloat16)
    dy = torch.randn(B * L, 1024, generator=g).cuda().to(torch.bfloat16)
    y, lse = lib.attn_fwd(qkv, B, L)
    got = lib.attn_bwd_fused(qkv, y, dy, lse, B, L)
    torch.cuda.synchronize()
    print('bwd fallback', bool(torch.isfinite(got.float()).all()))
    x = torch.randn(2, 128, 70, generator=g).cuda()  # latent block on the tensor cores (3xTF32 GEMMs + streaming kernels)
    w8 = [torch.ones(128), torch.randn(128, 1, 5, generator=g) * 0.3, torch.zeros(128), torch.randn(682, 128, 1, generator=g) * 0.09,
          torch.zeros(682), torch.randn(128, 341, 1, generator=g) * 0.05, torch.zeros(128), torch.ones(128)]
    w8 = [t.cuda() for t in w8]
    film = torch.randn(2, 384, generator=g).cuda() * 0.1
    yb = lib.lat_block_tc(x, w8, lib.lat_tc_pack(w8[3], w8[4], w8[5]), film)
    yr = lib.lat_block(x, w8, film)
    torch.cuda.synchronize()
    print('lat block tc vs fp32', float((yb - yr).abs().max()))
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

import torch
dev = torch.device("cuda", 0)
for kw in (dict(fused=True), dict(foreach=True), dict(foreach=False, fused=False)):
    p = [torch.nn.Parameter(torch.randn(8, device=dev)) for _ in range(2)]
    for q in p:
        q.grad = torch.randn(8, device=dev)
    o = torch.optim.AdamW(p, lr=1e-3, **kw)
    v0 = [q._version for q in p]
    o.step()
    print(kw, v0, [q._version for q in p])

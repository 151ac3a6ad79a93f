"""Seeded synthetic sample matrices and query sets of the benchmark shapes.
Used by bench.py and the parity tests only -- nothing here is on the product path.

  gauss    FP32 N(0,1): the worst case (distances concentrate near sqrt(2), top-k gaps ~1e-4)
  tissue   64 positive centroids, rows = centroid * (1 + 0.3 N(0,1)) under a fixed per-bucket sign:
           well-separated clusters, few ties
  sparse   intropolis-like: 1-3 non-zero buckets per row out of a small pool of junctions, integer
           coverage times one of a few idf values -- thousands of rows are parallel (distance-0 ties),
           as on tests/golden/tiny_intropolis.tsv (morna.py:344-388 builds such rows)
"""
import torch


def _gen(device, seed):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


def gauss(n, dim, device, seed=1234):
    return torch.randn((n, dim), generator=_gen(device, seed), device=device, dtype=torch.float32)


def tissue(n, dim, device, seed=1234, centroids=64):
    g = _gen(device, seed)
    c = torch.randn((centroids, dim), generator=g, device=device).abs() * torch.randn((centroids, dim), generator=g, device=device).exp()
    sign = torch.randint(0, 2, (dim,), generator=g, device=device).to(torch.float32) * 2 - 1
    t = torch.randint(0, centroids, (n,), generator=g, device=device)
    out = torch.empty((n, dim), device=device, dtype=torch.float32)
    step = 8192
    for r0 in range(0, n, step):                      # blockwise: no second n x dim temporary at 1 M rows
        r1 = min(n, r0 + step)
        noise = torch.randn((r1 - r0, dim), generator=g, device=device)
        out[r0:r1] = c[t[r0:r1]] * (1 + 0.3 * noise) * sign
    return out


def sparse(n, dim, device, seed=1234, junctions=40, max_nnz=3):
    """Rows with 1..max_nnz non-zeros drawn from `junctions` (bucket, sign, idf) triples."""
    g = _gen(device, seed)
    bucket = torch.randperm(dim, generator=g, device=device)[:junctions]
    sign = torch.randint(0, 2, (junctions,), generator=g, device=device).to(torch.float64) * 2 - 1
    idf = torch.rand((junctions,), generator=g, device=device, dtype=torch.float64) * 1.5 + 0.05
    out = torch.zeros((n, dim), device=device, dtype=torch.float32)
    nnz = torch.randint(1, max_nnz + 1, (n,), generator=g, device=device)
    rows = torch.arange(n, device=device)
    for j in range(max_nnz):
        pick = torch.randint(0, junctions, (n,), generator=g, device=device)
        cov = torch.randint(1, 6, (n,), generator=g, device=device).to(torch.float64)
        live = nnz > j
        val = (sign[pick] * (cov * idf[pick])).to(torch.float32)
        out[rows[live], bucket[pick[live]]] += val[live]
    return out


def sparse_fixture(n, dim, device, seed=1234):
    """Three junctions in all, as in tests/tiny_intropolis.tsv: about a third of the rows are parallel to any query."""
    return sparse(n, dim, device, seed, junctions=3, max_nnz=3)


KINDS = {"gauss": gauss, "tissue": tissue, "sparse": sparse, "sparse_fixture": sparse_fixture}


def matrix(kind, n, dim, device, seed=1234):
    return KINDS[kind](n, dim, device, seed)


def queries(S, nq, seed=99, noise=0.0):
    """Rows randperm(N, seed)[:nq] of S as float64 (in-index); `noise` > 0 adds noise * N(0,1) relative to the
    row's mean magnitude (out-of-index).  Returns (queries float64 [nq x dim], rows int64 [nq])."""
    n = S.shape[0]
    rows = torch.randperm(n, generator=torch.Generator().manual_seed(seed))[:nq].to(S.device)
    q = S[rows].to(torch.float64)
    if noise > 0:
        g = _gen(S.device, seed + 1)
        scale = q.abs().mean(dim=1, keepdim=True)
        q = q + noise * scale * torch.randn(q.shape, generator=g, device=S.device, dtype=torch.float64)
    return q, rows

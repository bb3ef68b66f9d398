"""The nearest-code search of csrc/distance.cu restated step by step in torch on the CPU (test infrastructure): bf16
candidate scores with fp32 accumulation, the two best per 32-code group, the threshold from the threads' pairs, the error
bound ``eps``, the overflow rule, the exact re-rank -- compared with the exhaustive search under the same exactly-defined
distance.  It checks the ARGUMENT that makes the kernel exact for any data (the bound really covers the bf16 error, the
overflow rule really catches hidden members) on adversarial banks; the kernel itself is checked on the GPU
(tests/test_gpu_parity.py::test_nearest_codes_exact_on_adversarial_banks)."""
import pytest
import torch

GROUP, THREADS = 32, 128          # kChunk, 32 * kRerankWarps


def exact_distance(X, Y):
    """fp64 reductions rounded once, the reference's association (YY + XX) - 2 YX, fp32 -- [n, m]."""
    Xd, Yd = X.double(), Y.double()
    xx = (Xd * Xd).sum(1).float()
    yy = (Yd * Yd).sum(1).float()
    yx = (Xd @ Yd.t()).float()
    return (yy[None, :] + xx[:, None]) - 2.0 * yx


def search(X, Y, k, eps_scale=1.0, stats=None):
    n, K = X.shape
    m = Y.shape[0]
    yy = (Y.double() ** 2).sum(1).float()
    xx = (X.double() ** 2).sum(1).float()
    dot = X.bfloat16().float() @ Y.bfloat16().float().t()                       # products exact in fp32, fp32 accumulation
    s_hat = yy[None, :] - 2.0 * dot                                             # [n, m]
    pad = (-m) % GROUP
    s_pad = torch.cat([s_hat, torch.full([n, pad], float('inf'))], 1).reshape(n, -1, GROUP)
    idx_pad = torch.arange(m + pad).reshape(-1, GROUP)
    two, pos = torch.topk(s_pad, 2, dim=2, largest=False, sorted=True)          # the GEMM epilogue's output: 2 per group
    D = exact_distance(X, Y)
    out_d, out_i = torch.empty([n, k]), torch.empty([n, k], dtype=torch.long)
    ymax = float(yy.max())
    for q in range(n):
        sc, ps = two[q], pos[q]                                                 # [groups, 2]
        ngroups = sc.shape[0]
        # pass 1: every thread's two smallest over its strided share of the groups; thr = 8th smallest of those pairs
        pairs = []
        for t in range(min(THREADS, ngroups)):
            mine = sc[t::THREADS].reshape(-1)
            pairs.append(torch.topk(mine, min(2, mine.numel()), largest=False).values)
        pool = torch.sort(torch.cat(pairs)).values
        pool = pool[torch.isfinite(pool)]
        thr = float(pool[7]) if pool.numel() >= 8 else float('inf')
        slack = max(1.02, 1.004 + K * 1.5259e-5)
        eps = eps_scale * (0.0078125 * slack * float(xx[q]) ** 0.5 * ymax ** 0.5 + 4.8e-7 * (float(xx[q]) + ymax))
        cut = thr + 2.0 * eps
        # pass 2: overflowed groups are rescanned, the candidates within the cut of the other groups survive
        ovf = sc[:, 1] <= cut
        ids = []
        for g in range(ngroups):
            if ovf[g]:
                ids += [int(v) for v in idx_pad[g] if v < m]
            else:
                ids += [int(idx_pad[g, ps[g, j]]) for j in range(2) if float(sc[g, j]) <= cut and int(idx_pad[g, ps[g, j]]) < m]
        if stats is not None:
            stats.append((len(ids), int(ovf.sum())))
        ids = torch.tensor(sorted(set(ids)), dtype=torch.long)
        d = D[q, ids]
        order = sorted(range(len(ids)), key=lambda j: (float(d[j]), int(ids[j])))[:k]
        out_d[q], out_i[q] = d[order], ids[order]
    return out_d, out_i


def exhaustive(X, Y, k):
    D = exact_distance(X, Y)
    d, i = torch.sort(D, dim=1, stable=True)                                    # ties to the lowest index
    return d[:, :k], i[:, :k]


def banks():
    g = torch.Generator().manual_seed(5)
    K = 64
    yield 'random', torch.randn([16, K], generator=g), torch.randn([700, K], generator=g)
    # clusters inside ONE group: 20 near-copies of the query's neighbour share a group, so the two kept candidates hide 18 others
    Y = torch.randn([640, K], generator=g)
    X = torch.randn([8, K], generator=g)
    for q in range(8):
        Y[64 * q + 3:64 * q + 23] = X[q] + 1e-3 * torch.randn([20, K], generator=g)
    yield 'clusters_in_one_group', X, Y
    # codes that differ by less than a bf16 ulp: identical candidate scores, the exact order decides
    base = torch.randn([1, K], generator=g)
    Y = base + 1e-4 * torch.randn([512, K], generator=g)
    yield 'sub_ulp_differences', base + 1e-4 * torch.randn([4, K], generator=g), Y
    # exact duplicates: ties go to the lowest index
    Y = torch.randn([96, K], generator=g).repeat(4, 1)
    yield 'duplicates', Y[:6] + 0.0, Y
    # tiny queries against large codes: all distances ~ |y|^2, decided in the last bits of the fp32 distance
    yield 'tiny_queries', 1e-3 * torch.randn([6, K], generator=g), 10.0 + torch.randn([300, K], generator=g)
    # a bank smaller than one group, and a ragged last group
    yield 'small_bank', torch.randn([5, K], generator=g), torch.randn([19, K], generator=g)
    # long codes (K beyond 1024: the fp32-accumulation term of the bound)
    yield 'long_codes', torch.randn([4, 2048], generator=g), torch.randn([200, 2048], generator=g)


@pytest.mark.parametrize('name,X,Y', list(banks()), ids=[b[0] for b in banks()])
def test_restated_search_equals_exhaustive_search(name, X, Y):
    k = 4
    d, i = search(X, Y, k)
    dr, ir = exhaustive(X, Y, k)
    assert torch.equal(i, ir), name
    assert torch.equal(d, dr), name


def test_the_bound_is_needed_and_the_margins_only_set_the_cost():
    """Without the margin (eps = 0) the bank of sub-ulp differences loses members (its bf16 scores are noise around one value:
    the exact best need not be among the 8 smallest of them) -- the bound is doing real work; with it, the number of exact
    evaluations on random data stays a small multiple of k."""
    bs = {b[0]: b for b in banks()}
    _, X, Y = bs['sub_ulp_differences']
    _, i0 = search(X, Y, 4, eps_scale=0.0)
    _, ir = exhaustive(X, Y, 4)
    assert not torch.equal(i0, ir)
    g = torch.Generator().manual_seed(9)
    X, Y = torch.randn([8, 512], generator=g), torch.randn([4096, 512], generator=g)
    stats = []
    _, i = search(X, Y, 4, stats=stats)
    assert torch.equal(i, exhaustive(X, Y, 4)[1])
    assert max(s[0] for s in stats) <= 400, stats

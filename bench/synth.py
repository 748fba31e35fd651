"""Synthetic read generators, shared by tests/ and bench.py.

numpy variant: small, CPU, deterministic -- used by parity tests (same bytes go to oracle and GPU).
torch variant: large, generated directly in HBM -- used by bench.py (inputs resident when timing starts).
"""
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def genome_codes(length, seed):
    return np.random.default_rng(seed).integers(0, 4, length, dtype=np.uint8)


def reads_numpy(n_reads, read_len=150, genome_len=100_000, seed=1, err=0.001, lowq=0.0005, n_rate=0.0, genome=None,
                var_len=False, starts=None):
    """Returns (bases u8[n*L], quals u8[...], off u64[n+1]).  Phred+33: 'I' everywhere, '+' at substitution errors,
    '#' (Q2 < minQuality) at a `lowq` fraction of bases, 'N' at an `n_rate` fraction."""
    rng = np.random.default_rng(seed + 77)
    g = genome if genome is not None else genome_codes(genome_len, seed)
    G = len(g)
    lens = np.full(n_reads, read_len, dtype=np.int64)
    if var_len:
        lens = rng.integers(max(1, read_len // 4), read_len + 1, n_reads)
    off = np.zeros(n_reads + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    total = int(off[-1])
    if starts is None:
        starts = rng.integers(0, G - read_len, n_reads)
    strand = rng.integers(0, 2, n_reads)
    codes = np.empty(total, dtype=np.uint8)
    if not var_len:
        # fixed length: gather whole blocks of reads at once (same bytes as the per-read loop below)
        ar = np.arange(read_len, dtype=np.int64)
        for c0 in range(0, n_reads, 200_000):
            c1 = min(n_reads, c0 + 200_000)
            seg = g[np.asarray(starts[c0:c1], dtype=np.int64)[:, None] + ar]
            rev = strand[c0:c1].astype(bool)
            seg[rev] = (3 - seg[rev])[:, ::-1]
            codes[c0 * read_len: c1 * read_len] = seg.reshape(-1)
    else:
        for i in range(n_reads):
            L = int(lens[i])
            seg = g[starts[i]: starts[i] + L]
            if strand[i]:
                seg = (3 - seg)[::-1]
            codes[int(off[i]): int(off[i]) + L] = seg
    quals = np.full(total, ord("I"), dtype=np.uint8)
    e = rng.random(total) < err
    codes[e] = (codes[e] + rng.integers(1, 4, int(e.sum()), dtype=np.uint8)) & 3
    quals[e] = ord("+")
    lq = rng.random(total) < lowq
    quals[lq] = ord("#")
    bases = ACGT[codes]
    if n_rate > 0:
        nn = rng.random(total) < n_rate
        bases = bases.copy()
        bases[nn] = ord("N")
    return np.ascontiguousarray(bases), quals, off


def metagenome_plan(n_genomes=1000, min_len=250_000, max_len=4_000_000, sigma=2.0, seed=0x4D455441):
    """Config C4: genome lengths log-uniform in [min_len, max_len], abundances log-normal(sigma); a read comes from genome i
    with probability ~ abundance_i * length_i.  Returns (lengths int64[n], read share float64[n])."""
    rng = np.random.default_rng(seed)
    lens = np.exp(rng.uniform(np.log(min_len), np.log(max_len), n_genomes)).astype(np.int64)
    ab = np.exp(rng.normal(0.0, sigma, n_genomes))
    share = ab * lens
    return lens, share / share.sum()


def reads_torch(n_reads, read_len=150, genome_len=250_000_000, seed=0x4B6D6572, err=0.001, lowq=0.0005, device="cuda",
                chunk=4_000_000, read_seed=0x5245414453, meta=None):
    """Config C2/C3-shaped reads generated in HBM.  Returns (bases u8[n*L], quals u8[n*L], off int64[n+1]) on `device`.
    meta = (lengths, share) from metagenome_plan: reads are drawn from many genomes with skewed abundances (C4)."""
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    g_off = g_len = cdf = None
    if meta is not None:
        lens, share = meta
        genome_len = int(lens.sum())
        g_off = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)).to(device)
        g_len = torch.from_numpy(lens.astype(np.int64)).to(device)
        cdf = torch.from_numpy(np.cumsum(share)).to(device)
    g = torch.randint(0, 4, (genome_len,), dtype=torch.uint8, device=device, generator=gen)
    gen.manual_seed(read_seed)            # same genome on every rank, rank-specific reads
    bases = torch.empty(n_reads * read_len, dtype=torch.uint8, device=device)
    quals = torch.empty(n_reads * read_len, dtype=torch.uint8, device=device)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    ar = torch.arange(read_len, device=device, dtype=torch.int64)
    for c0 in range(0, n_reads, chunk):
        n = min(chunk, n_reads - c0)
        if meta is None:
            starts = torch.randint(0, genome_len - read_len, (n,), device=device, generator=gen, dtype=torch.int64)
        else:
            gi = torch.searchsorted(cdf, torch.rand((n,), device=device, generator=gen, dtype=torch.float64)).clamp_(max=len(g_len) - 1)
            u = torch.rand((n,), device=device, generator=gen, dtype=torch.float64)
            starts = g_off[gi] + (u * (g_len[gi] - read_len).to(torch.float64)).to(torch.int64)
            del gi, u
        strand = torch.randint(0, 2, (n, 1), device=device, generator=gen, dtype=torch.int64).bool()
        idx_f = starts[:, None] + ar[None, :]
        idx_r = starts[:, None] + (read_len - 1 - ar)[None, :]
        codes = torch.where(strand, 3 - g[idx_r], g[idx_f])
        del idx_f, idx_r
        r = torch.rand((n, read_len), device=device, generator=gen)
        e = r < err
        shift = torch.randint(1, 4, (n, read_len), device=device, generator=gen, dtype=torch.uint8)
        codes = torch.where(e, (codes + shift) & 3, codes)
        q = torch.full((n, read_len), ord("I"), dtype=torch.uint8, device=device)
        q[e] = ord("+")
        q[(r >= err) & (r < err + lowq)] = ord("#")
        bases[c0 * read_len: (c0 + n) * read_len] = lut[codes.long()].reshape(-1)
        quals[c0 * read_len: (c0 + n) * read_len] = q.reshape(-1)
        del codes, r, e, shift, q
    off = torch.arange(n_reads + 1, device=device, dtype=torch.int64) * read_len
    del g
    return bases, quals, off

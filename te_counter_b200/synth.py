"""
Synthetic workloads of SURVEY.md 8(d): an "hg38-like" genes_tes index and BAM-derived read arrays
(bulk SE/PE, 10x-style single cell), seeded.  Used by bench.py and by the parity tests (the real
BAMs and the mm10/hg38 indices of BASELINE.json configs 1-3 are not available offline).

The index is built with numpy on the host; reads are generated with torch on whichever device is
asked for (cuda for the bench shapes, cpu for tests).
"""
import numpy as np

from .index import GlbIndex, T_GENE, T_TE

SEED = 20261018
HG38_LENGTHS = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636,
                138394717, 133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345,
                83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415, 16569]
HG38_NAMES = [str(i) for i in range(1, 23)] + ["X", "Y", "M"]


def synth_index(seed=SEED, n_te=4_600_000, n_exon=1_300_000, n_gene=38_000, n_te_names=1200,
                chrom_len=None, n_chrom=25):
    """TE features: length lognormal (median 300 bp), ~1200 `class:family:name` names drawn Zipf
    (s = 1.1), strand uniform.  Exons: 50-600 bp inside gene loci placed uniformly, type
    protein_coding 55 % / lncRNA 45 %, one strand per gene."""
    rng = np.random.default_rng(seed)
    if chrom_len is None:
        lens = np.array(HG38_LENGTHS[:n_chrom], dtype=np.int64)
        names_c = HG38_NAMES[:n_chrom]
    else:
        lens = np.full(n_chrom, chrom_len, dtype=np.int64)
        names_c = HG38_NAMES[:n_chrom]
    p = lens / lens.sum()
    # TEs
    te_chrom = rng.choice(n_chrom, size=n_te, p=p).astype(np.int32)
    te_len = np.clip(rng.lognormal(np.log(300.0), 0.8, size=n_te), 20, 8000).astype(np.int64)
    te_L = (rng.random(n_te) * np.maximum(lens[te_chrom] - te_len - 1, 1)).astype(np.int64)
    te_R = te_L + te_len
    classes = ["LINE", "LTR", "SINE", "DNA", "Retroposon", "tRNA"]
    te_names = ["%s:fam%d:rep%04d" % (classes[i % len(classes)], i % 53, i) for i in range(n_te_names)]
    w = 1.0 / np.arange(1, n_te_names + 1) ** 1.1
    te_name = rng.choice(n_te_names, size=n_te, p=w / w.sum())
    te_strand = rng.integers(0, 2, size=n_te).astype(np.uint8)
    # genes / exons
    g_chrom = rng.choice(n_chrom, size=n_gene, p=p).astype(np.int32)
    g_span = np.clip(rng.lognormal(np.log(30000.0), 1.0, size=n_gene), 2000, 1_000_000).astype(np.int64)
    g_span = np.minimum(g_span, np.maximum(lens[g_chrom] // 2, 700))
    g_start = (rng.random(n_gene) * np.maximum(lens[g_chrom] - g_span - 1, 1)).astype(np.int64)
    g_strand = rng.integers(0, 2, size=n_gene).astype(np.uint8)
    ex_gene = rng.integers(0, n_gene, size=n_exon)
    ex_len = rng.integers(50, 601, size=n_exon)
    ex_L = g_start[ex_gene] + (rng.random(n_exon) * np.maximum(g_span[ex_gene] - ex_len, 1)).astype(np.int64)
    ex_R = ex_L + ex_len
    gene_names = ["ENSG%011d" % i for i in range(n_gene)]
    names = sorted(set(te_names) | set(gene_names))
    nid = {k: i for i, k in enumerate(names)}
    te_ids = np.array([nid[k] for k in te_names], dtype=np.int32)
    gene_ids = np.array([nid[k] for k in gene_names], dtype=np.int32)
    chrom_id = np.concatenate([te_chrom, g_chrom[ex_gene]]).astype(np.int32)
    L = np.concatenate([te_L, ex_L]).astype(np.int32)
    R = np.concatenate([te_R, ex_R]).astype(np.int32)
    ensg = np.concatenate([te_ids[te_name], gene_ids[ex_gene]]).astype(np.int32)
    tcode = np.concatenate([np.full(n_te, T_TE, np.uint8), np.full(n_exon, T_GENE, np.uint8)])
    scode = np.concatenate([te_strand, g_strand[ex_gene]]).astype(np.uint8)
    idx = GlbIndex(names_c, chrom_id, L, R, ensg, tcode, scode, names)
    idx.chrom_lengths = lens
    return idx


def _torch():
    import torch
    return torch


def synth_bulk_reads(seed, idx, n_records, paired, device="cpu", edge_frac=0.001, sort=False, as_numpy=None,
                     shard=(0, 1)):
    """SURVEY.md 8(d) bulk reads.  paired: records come as adjacent mate pairs (name-collated),
    arrival order random.  shard=(rank, world): this rank's reads fall in its slice of the
    genome-linear coordinate axis (reads shard by genomic range across GPUs, BASELINE.json).
    Returns dict of arrays (numpy on cpu unless as_numpy=False)."""
    torch = _torch()
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    n_units = n_records // 2 if paired else n_records
    s = idx.sorted_layout()
    fL = torch.from_numpy(idx.L).to(dev)
    fC = torch.from_numpy(idx.chrom_id).to(dev)
    lens = torch.from_numpy(np.asarray(idx.chrom_lengths, dtype=np.int64)).to(dev)
    rnd = lambda n: torch.rand(n, generator=g, device=dev)
    rank, world = shard
    cum = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(lens, 0)])
    G = int(cum[-1].item())
    lin_lo, lin_hi = G * rank // world, G * (rank + 1) // world
    # features of this shard: contiguous in (chrom, L) order
    order = torch.from_numpy(s["order"]).to(dev)
    lin_sorted = cum[torch.from_numpy(idx.chrom_id[s["order"]]).to(dev).to(torch.int64)] + \
        torch.from_numpy(s["L"]).to(dev).to(torch.int64)
    f_lo = int(torch.searchsorted(lin_sorted, torch.tensor([lin_lo], device=dev)).item())
    f_hi = max(f_lo + 1, int(torch.searchsorted(lin_sorted, torch.tensor([lin_hi], device=dev)).item()))
    del lin_sorted
    pick = order[torch.randint(f_lo, min(f_hi, idx.n_features), (n_units,), generator=g, device=dev)]
    near = rnd(n_units) < 0.85
    c_near = fC[pick].to(torch.int64)
    s_near = fL[pick].to(torch.int64) + torch.randint(-100, 101, (n_units,), generator=g, device=dev)
    lin = lin_lo + (rnd(n_units).double() * float(lin_hi - lin_lo)).to(torch.int64)
    c_uni = (torch.searchsorted(cum, lin, right=True) - 1).clamp(0, idx.n_chrom - 1)
    s_uni = torch.minimum(lin - cum[c_uni], (lens[c_uni] - 200).clamp(min=0))
    del lin, order
    chrom = torch.where(near, c_near, c_uni)
    start = torch.where(near, s_near, s_uni).clamp(min=0)
    edge = rnd(n_units) < edge_frac
    e_start = (start // 10000) * 10000 + torch.where(rnd(n_units) < 0.5, 0, 9999)
    start = torch.where(edge, e_start, start)
    del pick, near, c_near, s_near, c_uni, s_uni, edge, e_start
    if paired:
        mate = (start + (250 + 50 * torch.randn(n_units, generator=g, device=dev)).to(torch.int64) - 100).clamp(min=0)
        st = torch.stack([start, mate], dim=1).reshape(-1)
        en = st + 100
        ch = torch.stack([chrom, chrom], dim=1).reshape(-1)
        del mate
    else:
        gap = rnd(n_units) < 0.15
        extra = torch.exp(7.0 + torch.randn(n_units, generator=g, device=dev)).to(torch.int64)
        st = start
        en = start + 100 + torch.where(gap, extra, torch.zeros_like(extra))
        ch = chrom
        del gap, extra
    del start, chrom
    n = st.numel()
    offc = rnd(n) < 0.01
    ch = torch.where(offc, torch.randint(idx.n_chrom, idx.n_chrom + 40, (n,), generator=g, device=dev), ch)
    r = rnd(n)
    mapq = torch.where(r < 0.8, torch.full((n,), 255, device=dev, dtype=torch.int64),
                       torch.where(r < 0.9, torch.randint(20, 60, (n,), generator=g, device=dev),
                                   torch.randint(0, 20, (n,), generator=g, device=dev)))
    flag = (rnd(n) < 0.01).to(torch.uint8) * 1 + (rnd(n) < 0.02).to(torch.uint8) * 2 + \
        (rnd(n) < 0.005).to(torch.uint8) * 4 + (rnd(n) < 0.5).to(torch.uint8) * 8
    if sort and not paired:
        key = ch * (1 << 32) + st
        order = torch.argsort(key)
        st, en, ch, mapq, flag = st[order], en[order], ch[order], mapq[order], flag[order]
    out = {"start": st.to(torch.int32), "end": en.clamp(max=2**31 - 16).to(torch.int32),
           "chrom": ch.to(torch.int32).to(torch.uint16) if hasattr(torch, "uint16") else ch.to(torch.int16),
           "mapq": mapq.to(torch.uint8), "flag": flag}
    if as_numpy is None:
        as_numpy = dev.type == "cpu"
    if as_numpy:
        out = {k: v.cpu().numpy() for k, v in out.items()}
        out["chrom"] = out["chrom"].view(np.uint16) if out["chrom"].dtype != np.uint16 else out["chrom"]
    return out


def synth_sc_reads(seed, idx, n_records, n_whitelist=100_000, n_cells=10_000, umis_per_cell=10_000,
                   umi_len=12, device="cpu", as_numpy=None, sort=True, part=(0, 1)):
    """SURVEY.md 8(d) single-cell reads: coordinate-sorted; `n_cells` real cells take 90 % of the
    reads (lognormal sizes), the other whitelist barcodes 8 %, 2 % not whitelisted; ~umis_per_cell
    distinct UMIs per real cell; 1 % of (cell, UMI) keys get a second chromosome/strand.
    part=(p, P): the p-th of P consecutive genome slices of one coordinate-sorted file -- every
    (cell, UMI) key lives in exactly one slice, so concatenating the parts gives the whole file."""
    torch = _torch()
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    n = n_records
    rnd = lambda k: torch.rand(k, generator=g, device=dev)
    fL = torch.from_numpy(idx.L).to(dev)
    fC = torch.from_numpy(idx.chrom_id).to(dev)
    n_cells = min(n_cells, n_whitelist)
    w = torch.exp(0.8 * torch.randn(n_cells, generator=g, device=dev)).double()
    real = torch.randperm(n_whitelist, generator=g, device=dev)[:n_cells]
    r = rnd(n)
    cell_real = real[torch.multinomial(w / w.sum(), n, replacement=True, generator=g)]
    cell_bg = torch.randint(0, n_whitelist, (n,), generator=g, device=dev)
    cell = torch.where(r < 0.90, cell_real, cell_bg)
    cell = torch.where(r >= 0.98, torch.full_like(cell, 0xFFFFFFFF), cell)
    del cell_real, cell_bg
    # UMI: a per-(cell, k) pseudo-random 12-mer over ACGT; k < umis_per_cell
    p_i, p_n = part
    if p_n > 1:
        g.manual_seed(int(seed) * 1009 + p_i)                    # same cells in every part, new reads
        k = torch.randint(0, max(1, umis_per_cell // p_n), (n,), generator=g, device=dev) * p_n + ((p_i - cell) % p_n)
    else:
        k = torch.randint(0, umis_per_cell, (n,), generator=g, device=dev)
    h = (cell * 1000003 + k * 7919 + 12345) & 0x7FFFFFFFFFFF
    h = (h * 6364136223846793005 + 1442695040888963407) & 0x7FFFFFFFFFFFFFFF
    code = torch.zeros(n, dtype=torch.int64, device=dev)
    lut = torch.tensor([1, 2, 3, 5], dtype=torch.int64, device=dev)      # A C G T (N = 4 unused)
    hh = h >> 8
    for i in range(umi_len):
        code = (code << 3) | lut[(hh >> (2 * i)) & 3]
    code = code << (3 * (21 - umi_len))
    # the read's place is a function of the key (a UMI tags one molecule) + small jitter; 1 % of
    # keys jump to a second place
    if p_n > 1:
        order = torch.from_numpy(idx.sorted_layout()["order"]).to(dev)
        f_lo, f_hi = idx.n_features * p_i // p_n, idx.n_features * (p_i + 1) // p_n
        span = max(1, f_hi - f_lo)
        hk = f_lo + (h >> 3) % span
        second = ((h >> 40) % 100 == 0) & (rnd(n) < 0.5)
        hk = order[torch.where(second, f_lo + ((hk - f_lo) * 31 + 17) % span, hk)]
        del order
    else:
        hk = (h >> 3) % idx.n_features
        second = ((h >> 40) % 100 == 0) & (rnd(n) < 0.5)
        hk = torch.where(second, (hk * 31 + 17) % idx.n_features, hk)
    chrom = fC[hk].to(torch.int64)
    start = (fL[hk].to(torch.int64) + torch.randint(-60, 61, (n,), generator=g, device=dev)).clamp(min=0)
    end = start + 91
    rev = ((h >> 50) & 1).to(torch.uint8)
    rev = torch.where(second, 1 - rev, rev)
    r2 = rnd(n)
    mapq = torch.where(r2 < 0.85, torch.full((n,), 255, device=dev, dtype=torch.int64),
                       torch.where(r2 < 0.93, torch.randint(20, 60, (n,), generator=g, device=dev),
                                   torch.randint(0, 20, (n,), generator=g, device=dev)))
    flag = (rnd(n) < 0.002).to(torch.uint8) * 1 + (rnd(n) < 0.01).to(torch.uint8) * 2 + \
        (rnd(n) < 0.002).to(torch.uint8) * 4 + rev * 8
    alt = rnd(n) < 0.005
    chrom = torch.where(alt, torch.full_like(chrom, 0xFFFE), chrom)
    if sort:
        order = torch.argsort(chrom * (1 << 32) + start)
        start, end, chrom, mapq, flag, cell, code = (t[order] for t in (start, end, chrom, mapq, flag, cell, code))
    out = {"start": start.to(torch.int32), "end": end.to(torch.int32), "chrom": chrom.to(torch.int32).to(torch.uint16),
           "mapq": mapq.to(torch.uint8), "flag": flag, "cell": cell.to(torch.uint32), "umi": code.to(torch.uint64)}
    if as_numpy is None:
        as_numpy = dev.type == "cpu"
    if as_numpy:
        out = {k: v.cpu().numpy() for k, v in out.items()}
    return out

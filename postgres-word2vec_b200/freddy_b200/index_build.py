"""Seeded synthetic index builder — stands in for the reference's
offline index_creation/*.py (scipy k-means + faiss flat search + SQL inserts),
which needs faiss/psycopg2/Postgres and is outside the search-parity boundary
("for a fixed index").  It follows the reference's *format and encode rule*:

  vectors   L2-normalised fp32 [N][d]            index_creation/index_utils.py:29-31
  coarse    C centroids, Lloyd k-means           quantizer_creation.py:31-33
  residual  m x K codewords on residuals         quantizer_creation.py:35-52
  codes     nearest codeword per sub-vector      ivfadc.py:60-90      (int16, ivfadc.py:112)
  ids       1-based, table order = id order      vec2database.py:84-87
  flat PQ   m x K codewords on raw sub-vectors   quantizer_creation.py:13-29, pq_index.py:106

torch is used as the array engine (GPU when available) — this is index
construction, not the search path.
"""
import numpy as np
import torch


def _kmeans(x, k, iters, gen, init=None):
    """Lloyd k-means on rows of x (float32 tensor); empty clusters keep their centroid.
    init: optional [k][d] starting centroids (default: k random rows)."""
    n = x.shape[0]
    perm = torch.randperm(n, generator=gen, device=x.device)[:k]
    cent = x[perm].clone() if init is None else init.clone()
    if k > n:  # tiny inputs: pad with jittered copies
        extra = x[torch.randint(0, n, (k - n,), generator=gen, device=x.device)]
        cent = torch.cat([cent, extra + 1e-3 * torch.randn(extra.shape, generator=gen, device=x.device)])
    for _ in range(iters):
        assign = _nearest(x, cent)
        sums = torch.zeros_like(cent).index_add_(0, assign, x)
        cnt = torch.zeros(k, device=x.device).index_add_(0, assign, torch.ones(n, device=x.device))
        nz = cnt > 0
        cent[nz] = sums[nz] / cnt[nz].unsqueeze(1)
    return cent


def _nearest(x, cent, chunk=65536):
    """argmin_j ||x_i - cent_j||^2, chunked (the reference uses faiss IndexFlatL2)."""
    out = torch.empty(x.shape[0], dtype=torch.long, device=x.device)
    cn = (cent * cent).sum(1)
    for s in range(0, x.shape[0], chunk):
        xs = x[s:s + chunk]
        d = cn.unsqueeze(0) - 2.0 * (xs @ cent.t())
        out[s:s + chunk] = d.argmin(1)
    return out


def make_synthetic_index(N, d=300, m=12, K=1024, C=1000, n_train=100_000, n_clusters=1000,
                         sigma=0.3, kmeans_iters=10, seed=1234, device=None, with_pq=False,
                         keep_vectors=False, zipf=0.7, coarse_from_centres=False, encoder=None):
    """Returns a dict of numpy arrays (plus 'vectors_t': the torch tensor, if keep_vectors).
    coarse_from_centres: start the coarse k-means from the (normalised) generating centres (needs C == n_clusters):
    one centroid per true cluster, i.e. inverted lists of the nominal N / C rows.
    encoder: a freddy_b200.Engine on the same CUDA device: coarse assignment and residual / pq codes of ALL rows then come from
    fb_encode_ivfadc_dev / fb_encode_pq_dev, i.e. from the reference's own rule (strict `<`, first minimum in table order,
    sequential fp32 distances; freddy.c:1567-1582, index_utils.c:923-939) instead of the ||c||^2 - 2 x.c shortcut below."""
    assert d % m == 0, "d must be divisible by m"
    sub = d // m
    dev = torch.device(device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu"))
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)

    # --- vectors: mixture of Gaussian clusters with Zipf-ish sizes, then normalised
    centres = torch.randn(n_clusters, d, generator=gen, device=dev)
    pc = 1.0 / torch.arange(10, 10 + n_clusters, device=dev, dtype=torch.float32) ** zipf
    pc = pc / pc.sum()
    vecs = torch.empty(N, d, device=dev, dtype=torch.float32)
    step = 262144
    for s in range(0, N, step):
        n = min(step, N - s)
        cl = torch.multinomial(pc, n, replacement=True, generator=gen)
        v = centres[cl] + sigma * torch.randn(n, d, generator=gen, device=dev)
        vecs[s:s + n] = v / v.norm(dim=1, keepdim=True)
    ntr = min(n_train, N)
    train = vecs[:ntr]

    # --- coarse quantizer
    init = None
    if coarse_from_centres and C == n_clusters:
        init = centres / centres.norm(dim=1, keepdim=True) / (1.0 + sigma * sigma) ** 0.5   # E[v | cluster] of the normalised rows
    coarse = _kmeans(train, C, kmeans_iters, gen, init=init)
    coarse_ids = _nearest(vecs, coarse)

    # --- residual PQ codebook (trained on the residuals of the training prefix)
    res_train = train - coarse[coarse_ids[:ntr]]
    res_cb = torch.empty(m, K, sub, device=dev)
    for p in range(m):
        res_cb[p] = _kmeans(res_train[:, p * sub:(p + 1) * sub].contiguous(), K, kmeans_iters, gen)
    codes = torch.empty(N, m, dtype=torch.int16, device=dev)
    encode_report = None
    if encoder is not None and dev.type == "cuda":
        from . import _lib
        encoder.load_coarse(coarse.cpu().numpy())
        encoder.load_codebook(_lib.FB_CB_RESIDUAL, res_cb.cpu().numpy())
        cids32 = torch.empty(N, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        encoder.encode_ivfadc_dev(vecs.data_ptr(), N, cids32.data_ptr(), codes.data_ptr())
        encoder.synchronize()
        # how far the float shortcut (argmin of ||c||^2 - 2 x.c) is from the reference's rule: rows it would assign / encode differently
        ns = min(N, 100_000)
        r = vecs[:ns] - coarse[cids32[:ns].to(torch.long)]
        short = torch.stack([_nearest(r[:, p * sub:(p + 1) * sub].contiguous(), res_cb[p]) for p in range(m)], 1).to(torch.int16)
        encode_report = {"rows": N, "coarse_ids_differing_from_the_float_shortcut": int((cids32.to(torch.long) != coarse_ids).sum().item()),
                         "sample_rows": ns, "sample_rows_with_other_codes": int((short != codes[:ns]).any(1).sum().item())}
        coarse_ids = cids32.to(torch.long)
    else:
        for s in range(0, N, step):
            r = vecs[s:s + step] - coarse[coarse_ids[s:s + step]]
            for p in range(m):
                codes[s:s + step, p] = _nearest(r[:, p * sub:(p + 1) * sub].contiguous(), res_cb[p]).to(torch.int16)

    out = {
        "d": d, "m": m, "K": K, "C": C, "N": N,
        "coarse": coarse.cpu().numpy(),
        "residual_codebook": res_cb.cpu().numpy(),
        "ids": np.arange(1, N + 1, dtype=np.int32),
        "coarse_ids": coarse_ids.to(torch.int32).cpu().numpy(),
        "codes": codes.cpu().numpy(),
    }
    if with_pq:
        pq_cb = torch.empty(m, K, sub, device=dev)
        for p in range(m):
            pq_cb[p] = _kmeans(train[:, p * sub:(p + 1) * sub].contiguous(), K, kmeans_iters, gen)
        pq_codes = torch.empty(N, m, dtype=torch.int16, device=dev)
        if encoder is not None and dev.type == "cuda":
            from . import _lib
            encoder.load_codebook(_lib.FB_CB_PQ, pq_cb.cpu().numpy())
            torch.cuda.synchronize()
            encoder.encode_pq_dev(vecs.data_ptr(), N, pq_codes.data_ptr())
            encoder.synchronize()
        else:
            for s in range(0, N, step):
                v = vecs[s:s + step]
                for p in range(m):
                    pq_codes[s:s + step, p] = _nearest(v[:, p * sub:(p + 1) * sub].contiguous(), pq_cb[p]).to(torch.int16)
        out["pq_codebook"] = pq_cb.cpu().numpy()
        out["pq_codes"] = pq_codes.cpu().numpy()
    if encode_report is not None:
        out["encode_report"] = encode_report
    if keep_vectors:
        out["vectors_t"] = vecs
    else:
        del vecs
    return out


def sample_queries(index, vectors_t, n, seed=4321):
    """Queries are stored vectors, as the reference's harness samples stored words
    (evaluation/evaluation_utils.py:103-116)."""
    g = torch.Generator()
    g.manual_seed(seed)
    sel = torch.randperm(vectors_t.shape[0], generator=g)[:n]
    return vectors_t[sel.to(vectors_t.device)].cpu().numpy(), (sel + 1).numpy().astype(np.int32)


def make_ivpq_index(vectors_t, m=12, K=1024, Kc=32, n_train=100_000, kmeans_iters=10, seed=77, target_rows=None):
    """IVPQ (inverted multi-index + PQ) tables for the kNN-join path, from normalised vectors:
      coarse_multi [2][Kc][d/2]   2-way product quantizer of the halves      ivpq.py:220-223
      codebook     [m][K][d/m]    PQ on the RAW vectors (not residuals)      ivpq.py:232-236
      coarse_ids   c0 + Kc*c1                                                 ivpq.py:18, :79
      stats        [Kc*Kc + 1]    relative cell frequency over `target_rows` (all rows if None),
                                  last entry = their count                    freddy--0.0.1.sql:150-171
    """
    dev = vectors_t.device
    N, d = vectors_t.shape
    assert d % 2 == 0 and d % m == 0
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    ntr = min(n_train, N)
    train = vectors_t[:ntr]
    half, sub = d // 2, d // m
    cm = torch.stack([_kmeans(train[:, h * half:(h + 1) * half].contiguous(), Kc, kmeans_iters, gen) for h in range(2)])
    cb = torch.stack([_kmeans(train[:, p * sub:(p + 1) * sub].contiguous(), K, kmeans_iters, gen) for p in range(m)])
    step = 262144
    cids = torch.empty(N, dtype=torch.int32, device=dev)
    codes = torch.empty(N, m, dtype=torch.int16, device=dev)
    for s in range(0, N, step):
        v = vectors_t[s:s + step]
        c0 = _nearest(v[:, :half].contiguous(), cm[0])
        c1 = _nearest(v[:, half:].contiguous(), cm[1])
        cids[s:s + step] = (c0 + Kc * c1).to(torch.int32)
        for p in range(m):
            codes[s:s + step, p] = _nearest(v[:, p * sub:(p + 1) * sub].contiguous(), cb[p]).to(torch.int16)
    sel = cids if target_rows is None else cids[torch.as_tensor(target_rows, device=dev, dtype=torch.long)]
    counts = torch.bincount(sel.to(torch.long), minlength=Kc * Kc).to(torch.float64)
    total = float(sel.numel())
    stats = torch.cat([(counts / total).to(torch.float32), torch.tensor([total], dtype=torch.float32, device=dev)])
    return {"d": d, "m": m, "K": K, "Kc": Kc, "N": N, "coarse_multi": cm.cpu().numpy(), "ivpq_codebook": cb.cpu().numpy(),
            "ids": np.arange(1, N + 1, dtype=np.int32), "ivpq_coarse_ids": cids.cpu().numpy(),
            "ivpq_codes": codes.cpu().numpy(), "stats": stats.cpu().numpy()}

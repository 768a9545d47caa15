"""vector_verify.py — a vectorised VERIFIER of QOA decode and encode results (torch, any device).

TEST INFRASTRUCTURE ONLY (see oracle/qoa_oracle.h): imported by tests/ alone, never by the product.

This is neither a decoder nor an encoder.  It takes a finished result and checks, for EVERY sample / slice at once,
that the result satisfies the reference's recurrences (paths into rafaelcaricio/qoaudio):

  decode (src/lib.rs:291-330, predict :799-805, update :821-828)
    With the claimed output x and the slice words known, the weight trajectory of a frame has the closed form
        w_k[i] = W_k + sum_{j<i} sgn(y[j+k]) * (dequantized_j >> 4),      y = header history ++ x,
    (one cumulative sum per tap, no sequential loop), and x is THE decode of the file iff
        x[i] == clamp(((sum_k w_k[i] * y[i+k]) wrapped to i32 >> 13) + dequantized_i, -32768, 32767)    for every i
    (induction over i: the first i that violates it is a wrong sample; no violation = the unique fixed point).
    Besides, the LMS state the decoder holds at the end of frame f must equal the state the ENCODER stored in the header
    of frame f+1 (:459-460 writes it, truncated to 16 bits): for a file written by somebody else's encoder — the
    reference's own fixture — that is a 128-bit value per channel and frame computed by that encoder over all 5120
    samples of the frame: a reference-held check of decoded VALUES, not only of counts.

  encode (encode_slice :495-596, qoa_div :613-617, predict_and_penalty_sq :808-818)
    The decode trajectory of the encoded stream gives the LMS state at the start of every slice (the encoder's own
    state as long as no weight left the 16-bit range at a frame boundary — :459-460 truncates what :482 carries in
    full; frames behind such a boundary are reported as unverifiable, they do not occur on music-like input).  From
    that state all 16 scalefactor candidates of ALL slices are run side by side (20 sequential steps over arrays of
    slices x 16), and the stored slice must be the candidate with the lexicographically smallest
    (total rank, first-sample rank, scalefactor) — what the reference's stable sort by first-sample rank (:535-541)
    followed by strict `<` replacement (:587) selects; pruning (:553, :578) cannot change the winner.

Formulation and language differ from oracle/qoa_oracle.c (sequential C, sorted + pruned search) and from
oracle/qoa_pyref.py (sequential Python): closed-form trajectories, slice-parallel search, 64-bit lanes with explicit
32-bit wrapping.  Works on CPU tensors (the -m "not gpu" tests) and on CUDA tensors (full-size checks of the GPU
kernels' output with torch's own elementwise kernels: nothing of the product is involved in the checking).

Scope: batches of equally shaped conforming streams ([S, stride] bytes; every frame but the last has 5120 samples).
"""
import torch

SLICE_LEN = 20
FRAME_LEN = 5120
SLICES_PER_FRAME = 256
MAGIC = 0x716F6166

_QUANT = [7, 7, 7, 5, 5, 3, 3, 1, 0, 0, 2, 2, 4, 4, 6, 6, 6]                                         # src/lib.rs:22
_RECIP = [65536, 9363, 3121, 1457, 781, 475, 311, 216, 156, 117, 90, 71, 57, 47, 39, 32]             # :25-27
_MAG = [(1, 3, 5, 7), (5, 18, 32, 49), (16, 53, 95, 147), (34, 113, 203, 315), (63, 210, 378, 588),  # :847-864
        (104, 345, 621, 966), (158, 528, 950, 1477), (228, 760, 1368, 2128), (316, 1053, 1895, 2947),
        (422, 1405, 2529, 3934), (548, 1828, 3290, 5117), (696, 2320, 4176, 6496), (868, 2893, 5207, 8099),
        (1064, 3548, 6386, 9933), (1286, 4288, 7718, 12005), (1536, 5120, 9216, 14336)]
_DEQ = [[(-_MAG[sf][q >> 1] if q & 1 else _MAG[sf][q >> 1]) for q in range(8)] for sf in range(16)]


def encoded_size(ch, n):  # :374-378
    return 8 + ((n + FRAME_LEN - 1) // FRAME_LEN) * (8 + 16 * ch) + ((n + SLICE_LEN - 1) // SLICE_LEN) * 8 * ch


def _wrap32(v):
    return ((v + 0x80000000) & 0xFFFFFFFF) - 0x80000000


def _wrap16(v):
    return ((v + 0x8000) & 0xFFFF) - 0x8000


def _sign(v):
    return (v > 0).to(torch.int64) - (v < 0).to(torch.int64)


def _halves(b):
    """[..., 8k] bytes -> (hi, lo): the big-endian 64-bit words as two non-negative int64 halves, each [..., k]."""
    b = b.reshape(*b.shape[:-1], -1, 8).to(torch.int64)
    hi = (b[..., 0] << 24) | (b[..., 1] << 16) | (b[..., 2] << 8) | b[..., 3]
    lo = (b[..., 4] << 24) | (b[..., 5] << 16) | (b[..., 6] << 8) | b[..., 7]
    return hi, lo


def _lms_fields(hi, lo):
    """a history / weights word -> [..., 4] sign-extended 16-bit fields (:271-281)"""
    return _wrap16(torch.stack([hi >> 16, hi & 0xFFFF, lo >> 16, lo & 0xFFFF], dim=-1))


def _codes(hi, lo):
    """slice words -> (sf [...], codes [..., 20]): sf = bits 63..60, sample i = bits 59-3i .. 57-3i (:303, :308)"""
    out = []
    for i in range(SLICE_LEN):
        s = 57 - 3 * i
        if s >= 32:
            out.append((hi >> (s - 32)) & 7)
        elif s == 30:  # bits 32..30 straddle the halves
            out.append(((hi & 1) << 2) | (lo >> 30))
        else:
            out.append((lo >> s) & 7)
    return hi >> 28, torch.stack(out, dim=-1)


class _Frames:
    """One batch of equally long frames: B frames x ch channels x K slices.  Holds the closed-form trajectory."""

    def __init__(self, H, W, sf, codes, x):
        # H, W [B, ch, 4]; sf [B, K, ch]; codes [B, K, ch, 20]; x [B, ch, K * 20] claimed samples (int64)
        dev = x.device
        B, ch, n = x.shape
        deq = torch.tensor(_DEQ, dtype=torch.int64, device=dev)
        dq = deq[sf.unsqueeze(-1).expand_as(codes), codes]               # [B, K, ch, 20]
        dq = dq.permute(0, 2, 1, 3).reshape(B, ch, n)
        delta = dq >> 4                                                    # :822
        y = torch.cat([H, x], dim=2)                                       # history ++ samples
        acc = torch.zeros_like(x)
        w_before, w_end = [], []
        for k in range(4):
            yk = y[:, :, k:k + n]
            upd = torch.where(yk < 0, -delta, delta)                       # :823-826 (zero counts as positive)
            cs = torch.cumsum(upd, dim=2)
            wk = _wrap32(W[:, :, k:k + 1] + cs - upd)                      # weights in front of step i
            acc = acc + wk * yk
            w_before.append(wk[:, :, ::SLICE_LEN])                         # ... in front of every slice
            w_end.append(_wrap32(W[:, :, k] + cs[:, :, -1]))
        p = _wrap32(acc) >> 13                                             # :799-805
        self.bad = torch.clamp(p + dq, -32768, 32767) != x                 # :310
        self.w_end = torch.stack(w_end, dim=-1)                            # [B, ch, 4]
        self.h_end = y[:, :, n:n + 4]
        self.w_slice = torch.stack(w_before, dim=-1)                       # [B, ch, K, 4]
        self.h_slice = torch.stack([y[:, :, k:k + n:SLICE_LEN] for k in range(4)], dim=-1)


def _parse(qoa, ch, n):
    """[S, >= size] bytes of S equally shaped conforming streams -> header words and per-frame-group fields."""
    S = qoa.shape[0]
    size = encoded_size(ch, n)
    hi, lo = _halves(qoa[:, :size])
    full, tail = n // FRAME_LEN, n % FRAME_LEN
    fw = 1 + 2 * ch + SLICES_PER_FRAME * ch  # words per full frame
    groups = []
    if full:
        fh, fl = hi[:, 1:1 + full * fw].reshape(S, full, fw), lo[:, 1:1 + full * fw].reshape(S, full, fw)
        groups.append((fh, fl, SLICES_PER_FRAME, FRAME_LEN))
    if tail:
        k = (tail + SLICE_LEN - 1) // SLICE_LEN
        groups.append((hi[:, 1 + full * fw:].reshape(S, 1, -1), lo[:, 1 + full * fw:].reshape(S, 1, -1), k, tail))
    out = []
    for gh, gl, k, nspc in groups:
        assert gh.shape[2] == 1 + 2 * ch + k * ch
        hdr = (gh[:, :, 0], gl[:, :, 0])
        lh, ll = gh[:, :, 1:1 + 2 * ch].reshape(S, -1, ch, 2), gl[:, :, 1:1 + 2 * ch].reshape(S, -1, ch, 2)
        H = _lms_fields(lh[..., 0], ll[..., 0])                            # [S, F, ch, 4]
        W = _lms_fields(lh[..., 1], ll[..., 1])
        sh, sl = gh[:, :, 1 + 2 * ch:].reshape(S, -1, k, ch), gl[:, :, 1 + 2 * ch:].reshape(S, -1, k, ch)
        sf, codes = _codes(sh, sl)
        out.append(dict(hdr=hdr, H=H, W=W, sf=sf, codes=codes, k=k, nspc=nspc))
    return (hi[:, 0], lo[:, 0]), out


def _group_samples(pcm, ch, n, first_sample, k, nspc, frames):
    """[S, n * ch] interleaved -> [S * frames, ch, k * 20] int64 (zero-padded behind a partial last slice)"""
    S = pcm.shape[0]
    x = pcm[:, first_sample * ch:(first_sample + frames * nspc) * ch].reshape(S * frames, nspc, ch).permute(0, 2, 1)
    x = x.to(torch.int64)
    pad = k * SLICE_LEN - nspc
    if pad:
        x = torch.nn.functional.pad(x, (0, pad))
    return x.contiguous()


def verify_decode(qoa, pcm, ch, n, rate=None, want_trajectory=False):
    """qoa [S, >= encoded_size] uint8, pcm [S, n * ch] int16 (the claimed decode of every stream).
    Returns a dict of counts (all zero = every sample of every stream is the reference's decode of its file)."""
    S = qoa.shape[0]
    (f_hi, f_lo), groups = _parse(qoa, ch, n)
    res = dict(streams=S, samples=S * n * ch, bad_file_headers=int(((f_hi != MAGIC) | (f_lo != n)).sum()),
               bad_frame_headers=0, bad_samples=0, boundaries=0, chain_history_bad=0, chain_weights_bad=0,
               truncated_boundaries=0)
    bad_streams = torch.zeros(S, dtype=torch.bool, device=qoa.device)
    chain_bad_streams = torch.zeros(S, dtype=torch.bool, device=qoa.device)
    first, prev_end, traj = 0, None, []
    for g in groups:
        F, k, nspc = g["H"].shape[1], g["k"], g["nspc"]
        exp_lo = (nspc << 16) | ((8 + 16 * ch + 8 * k * ch) & 0xFFFF)     # :449-452
        bad_h = (g["hdr"][1] != exp_lo) | ((g["hdr"][0] >> 24) != ch)
        if rate is not None:
            bad_h |= (g["hdr"][0] & 0xFFFFFF) != rate
        res["bad_frame_headers"] += int(bad_h.sum())
        x = _group_samples(pcm, ch, n, first, k, nspc, F)
        fr = _Frames(g["H"].reshape(S * F, ch, 4), g["W"].reshape(S * F, ch, 4), g["sf"].reshape(S * F, k, ch),
                     g["codes"].reshape(S * F, k, ch, SLICE_LEN), x)
        bad = fr.bad[:, :, :nspc]
        res["bad_samples"] += int(bad.sum())
        bad_streams |= bad.reshape(S, -1).any(dim=1)
        # chain: state at the end of frame f (full frames only) against the header of frame f + 1
        h_end, w_end = fr.h_end.reshape(S, F, ch, 4), fr.w_end.reshape(S, F, ch, 4)
        nxt_H, nxt_W = g["H"][:, 1:], g["W"][:, 1:]
        pairs = [(h_end[:, :-1], w_end[:, :-1], nxt_H, nxt_W)] if F > 1 else []
        if prev_end is not None:
            pairs.append((prev_end[0], prev_end[1], g["H"][:, :1], g["W"][:, :1]))
        for he, we, nh, nw in pairs:
            res["boundaries"] += int(he.shape[0] * he.shape[1] * ch)
            hb, wb = (he != nh).any(dim=-1), (_wrap16(we) != nw).any(dim=-1)   # [S, boundaries, ch]
            res["chain_history_bad"] += int(hb.sum())
            res["chain_weights_bad"] += int(wb.sum())
            chain_bad_streams |= (hb | wb).reshape(S, -1).any(dim=1)
            res["truncated_boundaries"] += int((_wrap16(we) != we).any(dim=-1).sum())
        prev_end = (h_end[:, -1:], w_end[:, -1:]) if nspc == FRAME_LEN else None
        if want_trajectory:
            traj.append((g, fr, first, F))
        first += F * nspc
    res["bad_streams"] = int(bad_streams.sum())
    res["chain_bad_streams"] = int(chain_bad_streams.sum())
    if want_trajectory:
        res["_trajectory"] = traj
    return res


_INIT_W = [0, 0, -(1 << 13), 1 << 14]  # :348-351


def verify_encode(qoa, src, dec, ch, n, rate=None, max_lanes=1 << 24):
    """qoa [S, >= encoded_size] uint8 = the encoder's output for src [S, n * ch] int16; dec [S, n * ch] int16 = a decode of
    qoa (checked here with verify_decode before it is used).  Returns counts; all zero = every slice of every stream is
    what the reference's search selects, every frame header carries the encoder's state."""
    dev = qoa.device
    S = qoa.shape[0]
    res = verify_decode(qoa, dec, ch, n, rate, want_trajectory=True)
    traj = res.pop("_trajectory")
    res.update(slices=0, bad_slices=0, unverifiable_slices=0, bad_initial_state=0)
    quant = torch.tensor(_QUANT, dtype=torch.int64, device=dev)
    deq = torch.tensor(_DEQ, dtype=torch.int64, device=dev)
    recip = torch.tensor(_RECIP, dtype=torch.int64, device=dev).unsqueeze(0)
    sfv = torch.arange(16, dtype=torch.int64, device=dev).unsqueeze(0)
    init_w = torch.tensor(_INIT_W, dtype=torch.int64, device=dev)
    carried_ok = torch.ones(S, ch, dtype=torch.bool, device=dev)  # the decoder's state still equals the encoder's
    for gi, (g, fr, first, F) in enumerate(traj):
        k, nspc = g["k"], g["nspc"]
        if gi == 0:  # the first frame header holds the initial state (:348-351)
            res["bad_initial_state"] += int(((g["H"][:, 0] != 0).any(dim=-1) | (g["W"][:, 0] != init_w).any(dim=-1)).sum())
        # Frames whose header lost weight bits (:459-460) start from a state that is not the encoder's any more.
        w_end = fr.w_end.reshape(S, F, ch, 4)
        lost = (_wrap16(w_end) != w_end).any(dim=-1)                       # [S, F, ch]: boundary BEHIND frame f
        ok_f = torch.ones(S, F, ch, dtype=torch.bool, device=dev)
        ok_f[:, 0] = carried_ok
        if F > 1:
            ok_f[:, 1:] = ~lost[:, :-1]
            ok_f = torch.cumprod(ok_f.to(torch.int64), dim=1).bool()
        carried_ok = ok_f[:, -1] & ~lost[:, -1]
        s_all = _group_samples(src, ch, n, first, k, nspc, F).reshape(S * F, ch, k, SLICE_LEN)
        # slices as a flat list of lanes: [N, ...]
        h0 = fr.h_slice.reshape(-1, 4)
        w0 = fr.w_slice.reshape(-1, 4)
        smp = s_all.reshape(-1, SLICE_LEN)
        st_sf = g["sf"].reshape(S * F, k, ch).permute(0, 2, 1).reshape(-1)
        st_codes = g["codes"].reshape(S * F, k, ch, SLICE_LEN).permute(0, 2, 1, 3).reshape(-1, SLICE_LEN)
        lens = torch.full((S * F, ch, k), SLICE_LEN, dtype=torch.int64, device=dev)
        if nspc % SLICE_LEN:
            lens[:, :, -1] = nspc % SLICE_LEN
        lens = lens.reshape(-1)
        usable = ok_f.reshape(S * F, ch, 1).expand(S * F, ch, k).reshape(-1)
        N = h0.shape[0]
        res["slices"] += N
        res["unverifiable_slices"] += int((~usable).sum())
        step = max(1, max_lanes // 16)
        for b in range(0, N, step):
            e = min(N, b + step)
            bad = _search_mismatch(h0[b:e], w0[b:e], smp[b:e], lens[b:e], st_sf[b:e], st_codes[b:e], quant, deq, recip, sfv)
            res["bad_slices"] += int((bad & usable[b:e]).sum())
    return res


def _search_mismatch(h, w, smp, lens, st_sf, st_codes, quant, deq, recip, sfv):
    """[M] slices: run all 16 candidates from state (h, w) over smp[:, :len]; True where the stored (sf, codes) is not
    the lexicographic minimum of (total rank, first-sample rank, sf)."""
    M = h.shape[0]
    hs = [h[:, k:k + 1].expand(M, 16).clone() for k in range(4)]
    ws = [w[:, k:k + 1].expand(M, 16).clone() for k in range(4)]
    total = torch.zeros(M, 16, dtype=torch.int64, device=h.device)
    first = None
    codes = []
    sf_idx = sfv.expand(M, 16)
    for i in range(SLICE_LEN):
        act = (lens > i).unsqueeze(1)
        s = smp[:, i:i + 1]
        p = _wrap32(ws[0] * hs[0] + ws[1] * hs[1] + ws[2] * hs[2] + ws[3] * hs[3]) >> 13          # :808-812
        pen = torch.clamp_min((_wrap32(ws[0] * ws[0] + ws[1] * ws[1] + ws[2] * ws[2] + ws[3] * ws[3]) >> 18) - 0x8FF, 0)
        v = s - p                                                                                   # :567
        nn = _wrap32(_wrap32(v * recip) + (1 << 15)) >> 16                                          # qoa_div :613-617
        scaled = nn + _sign(v) - _sign(nn)
        q = quant[torch.clamp(scaled, -8, 8) + 8]                                                  # :568-571
        dq = deq[sf_idx, q]
        r = torch.clamp(p + dq, -32768, 32767)                                                     # :573
        err = s - r
        rank = err * err + pen * pen                                                               # :575-576
        if i == 0:
            first = rank
        total = total + torch.where(act, rank, torch.zeros_like(rank))
        d = dq >> 4                                                                                # update :821-828
        for k in range(4):
            ws[k] = torch.where(act, _wrap32(ws[k] + torch.where(hs[k] < 0, -d, d)), ws[k])
        hs = [torch.where(act, hs[1], hs[0]), torch.where(act, hs[2], hs[1]), torch.where(act, hs[3], hs[2]),
              torch.where(act, r, hs[3])]
        codes.append(torch.where(act, q, torch.zeros_like(q)))                                     # :485-487: zero-filled tail
    big = torch.iinfo(torch.int64).max
    t_min = total.min(dim=1, keepdim=True).values
    tie = total == t_min
    f_min = torch.where(tie, first, torch.full_like(first, big)).min(dim=1, keepdim=True).values
    tie &= first == f_min
    win = torch.where(tie, sf_idx, torch.full_like(sf_idx, 99)).min(dim=1).values                  # lowest sf among the ties
    cand = torch.stack(codes, dim=-1)                                                              # [M, 16, 20]
    win_codes = cand[torch.arange(M, device=h.device), win]
    return (win != st_sf) | (win_codes != st_codes).any(dim=1)

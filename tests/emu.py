"""NumPy emulation of the denoiser kernels' arithmetic, operand layouts and index math.

Test infrastructure: mirrors what csrc/denoiser.cu does tile by tile (same CP8 planes,
same tap shifts, same guards, same packed weights from prior_diffuse_b200/pack.py) in
float64, so that layout / packing / shift mistakes are caught on CPU against the oracle
before any GPU time is spent.  Variable names follow the kernels.
"""
import numpy as np

from prior_diffuse_b200 import pack as P


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def prelu(x, a):
    return np.where(x >= 0, x, a * x)


# ------------------------------------------------------------------ CP8 "split" layout
def split_q(F):
    return (F + 1) // 2


def to_cp8_split(x):
    """[B,C,T,F] -> [B][C/8][T*2Q][8]; pos(t,f) = t*2Q + (f&1)*Q + (f>>1)."""
    B, C, T, F = x.shape
    Q = split_q(F)
    out = np.zeros((B, C // 8, T * 2 * Q, 8))
    for f in range(F):
        pos = np.arange(T) * 2 * Q + (f & 1) * Q + (f >> 1)
        out[:, :, pos, :] = x[:, :, :, f].reshape(B, C // 8, 8, T).transpose(0, 1, 3, 2)
    return out


def from_cp8_split(a, F):
    B, CC, NP, _ = a.shape
    Q = split_q(F)
    T = NP // (2 * Q)
    out = np.zeros((B, CC * 8, T, F))
    for f in range(F):
        pos = np.arange(T) * 2 * Q + (f & 1) * Q + (f >> 1)
        out[:, :, :, f] = a[:, :, pos, :].transpose(0, 1, 3, 2).reshape(B, CC * 8, T)
    return out


def gemm_planes(A, row0, nrows, W):
    """D[m][n] = sum_{kc,j} A[kc][row0+m][j] * W[kc][n][j]  (A window shifted by row0)."""
    return np.einsum("kmj,knj->mn", A[:, row0:row0 + nrows, :], W)


# ------------------------------------------------------------------ time embedding
def emu_time(tp, t):
    """t: float array [B] -> bias rows [B][452]  (pdse_time_embed)."""
    t = np.asarray(t, dtype=np.float64)
    lo, hi = np.floor(t).astype(int), np.ceil(t).astype(int)
    tab = tp["table"].astype(np.float64)
    e = tab[lo] + (tab[hi] - tab[lo]) * (t - lo)[:, None]
    h = e @ tp["p1w"].T.astype(np.float64) + tp["p1b"]
    h = h * sigmoid(h)
    h = h @ tp["p2w"].T.astype(np.float64) + tp["p2b"]
    h = h * sigmoid(h)
    return h @ tp["rows"].T.astype(np.float64) + tp["bias"]


# ------------------------------------------------------------------ GLU tail
def glu_tail(blob, D2):
    """D2 [M][64] = l|r pre-bias accumulators -> block output [M][64] (or [M] for de1)."""
    lr = D2 + blob.f["blr"]
    l, r = lr[:, :32], lr[:, 32:]
    A2 = lr.reshape(-1, 8, 8).transpose(1, 0, 2)                  # [8][M][8]
    D3l = gemm_planes(A2[0:4], 0, A2.shape[1], blob.h["wgl"])
    D3r = gemm_planes(A2[4:8], 0, A2.shape[1], blob.h["wgr"])
    lm = sigmoid(D3l + blob.f["bg"][:32])
    rm = sigmoid(D3r + blob.f["bg"][32:])
    g = l * rm + r * lm
    if "w2" in blob.h:
        A3 = g.reshape(-1, 4, 8).transpose(1, 0, 2)
        D4 = gemm_planes(A3, 0, A3.shape[1], blob.h["w2"])
        return prelu(D4 * blob.f["scale"] + blob.f["shift"], blob.f["slope"][0])
    return g @ blob.f["w2vec"] + blob.f["b2"][0]


# ------------------------------------------------------------------ encoder block 1
def emu_enc1(blob, x, x0, bias_rows):
    """x, x0 [B,2,T,161] fp -> e1 CP8 split (F=79).  Tile = 128 consecutive output positions."""
    B, _, T, F = x.shape
    Fo, Qo = 79, 40
    wp, bp = blob.f["wp"].reshape(2, 4), blob.f["bp"][:2]
    out = np.zeros((B, 8, T * 2 * Qo, 8))
    for b in range(B):
        tb = bias_rows[b][0:2]
        cat = np.concatenate([x[b], x0[b]], axis=0)              # [4][T][F]
        u = np.einsum("cj,jtf->ctf", wp, cat) + bp[:, None, None] + tb[:, None, None]
        u = np.concatenate([np.broadcast_to(tb[:, None, None], (2, 1, F)), u], axis=1)   # row 0 = pad row (t=-1)
        npos = T * 2 * Qo
        for p0 in range(0, npos, 128):
            A = np.zeros((128, 32))
            meta = []
            for m in range(128):
                p = p0 + m
                t, rem = p // (2 * Qo), p % (2 * Qo)
                par, q = rem // Qo, rem % Qo
                fo = 2 * q + par
                ok = p < npos and fo < Fo
                meta.append((p, ok))
                if not ok:
                    continue
                for c in range(2):
                    for dt in range(2):
                        for df in range(5):
                            A[m, c * 10 + dt * 5 + df] = u[c, t + dt, 2 * fo + df]   # u row index t+dt <-> time t-1+dt
            Ap = A.reshape(128, 4, 8).transpose(1, 0, 2)
            D2 = gemm_planes(Ap, 0, 128, blob.h["wf"])
            y = glu_tail(blob, D2)
            for m, (p, ok) in enumerate(meta):
                if p < npos:
                    out[b, :, p, :] = y[m].reshape(8, 8) if ok else 0.0
    return out


# ------------------------------------------------------------------ encoder blocks 2..5
def emu_enc(blob, xin, Fin, hb_rows, nt):
    """xin CP8 split [B][8][T*2Qi][8] -> out CP8 split (Fout = (Fin-3)//2+1)."""
    B = xin.shape[0]
    Qi = split_q(Fin)
    T = xin.shape[2] // (2 * Qi)
    Fo = (Fin - 3) // 2 + 1
    Qo = split_q(Fo)
    Pp = Qi                                   # virtual row pitch
    out = np.zeros((B, 8, T * 2 * Qo, 8))
    HP = (nt + 1) * Pp + 2 + 128              # plane length incl. read slack
    for b in range(B):
        hb = hb_rows[b]
        for t0 in range(0, T, nt):
            # X planes: time rows t0-1 .. t0+nt-1
            X = np.zeros((8, (nt + 1) * 2 * Qi + 128, 8))
            for tl in range(nt + 1):
                t = t0 - 1 + tl
                if 0 <= t < T:
                    X[:, tl * 2 * Qi:(tl + 1) * 2 * Qi] = xin[b, :, t * 2 * Qi:(t + 1) * 2 * Qi]
            rows1 = (nt + 1) * 2 * Qi
            H = np.zeros((4, 2, HP, 8))
            for r0 in range(0, rows1, 128):
                D1 = gemm_planes(X, r0, 128, blob.h["w1"])
                for m in range(128):
                    r = r0 + m
                    if r >= rows1:
                        continue
                    tl, rem = r // (2 * Qi), r % (2 * Qi)
                    par, q = rem // Qi, rem % Qi
                    H[:, par, tl * Pp + q, :] = (D1[m] + hb).reshape(4, 8)
            Hf = H.reshape(8, HP, 8)          # plane index = cc*2 + par
            for m0 in range(0, nt * Pp, 128):
                D2 = np.zeros((128, 64))
                for dt in range(2):
                    for df in range(3):
                        par, sh = df & 1, dt * Pp + (df >> 1)
                        Aw = Hf[par::2]       # [4][HP][8] planes of this parity
                        D2 += gemm_planes(Aw, m0 + sh, 128, blob.h["wlr"][dt * 3 + df])
                y = glu_tail(blob, D2)
                for m in range(128):
                    mm = m0 + m
                    tl, j = mm // Pp, mm % Pp
                    t = t0 + tl
                    if tl < nt and j < Fo and t < T:
                        out[b, :, t * 2 * Qo + (j & 1) * Qo + (j >> 1), :] = y[m].reshape(8, 8)
    return out


# ------------------------------------------------------------------ decoder blocks
def emu_dec(blob, xa, skip, Fin, kw, hb_rows, nt):
    """xa, skip CP8 split [B][8][T*2Qi][8] -> out CP8 split (Fout = 2Fin+kw-2) or eps [B][T][Fout]."""
    B = xa.shape[0]
    Qi = split_q(Fin)
    T = xa.shape[2] // (2 * Qi)
    G = (kw - 1) // 2
    Pp = Fin + G
    Fo = 2 * Fin + kw - 2
    last = "w2" not in blob.h
    out = np.zeros((B, T, Fo)) if last else np.zeros((B, 8, T * 2 * Pp, 8))
    HP = (nt + 1) * Pp + G + 128
    for b in range(B):
        hb = hb_rows[b]
        for t0 in range(0, T, nt):
            X = np.zeros((16, (nt + 1) * 2 * Qi + 128, 8))
            for tl in range(nt + 1):
                t = t0 - 1 + tl
                if 0 <= t < T:
                    X[0:8, tl * 2 * Qi:(tl + 1) * 2 * Qi] = xa[b, :, t * 2 * Qi:(t + 1) * 2 * Qi]
                    X[8:16, tl * 2 * Qi:(tl + 1) * 2 * Qi] = skip[b, :, t * 2 * Qi:(t + 1) * 2 * Qi]
            rows1 = (nt + 1) * 2 * Qi
            H = np.zeros((4, HP, 8))          # guards stay zero
            for r0 in range(0, rows1, 128):
                D1 = gemm_planes(X, r0, 128, blob.h["w1"])
                for m in range(128):
                    r = r0 + m
                    if r >= rows1:
                        continue
                    tl, rem = r // (2 * Qi), r % (2 * Qi)
                    par, q = rem // Qi, rem % Qi
                    f = 2 * q + par
                    t = t0 - 1 + tl
                    if f < Fin and 0 <= t < T:
                        H[:, tl * Pp + f + G, :] = (D1[m] + hb).reshape(4, 8)
            for parity in range(2):
                taps = [(dt, a) for dt in range(2) for a in range(G + 1 - parity)]
                W = blob.h["wlr_even" if parity == 0 else "wlr_odd"]
                for m0 in range(0, nt * Pp, 128):
                    D2 = np.zeros((128, 64))
                    for ti, (dt, a) in enumerate(taps):
                        sh = (1 - dt) * Pp + G - a
                        D2 += gemm_planes(H, m0 + sh, 128, W[ti])
                    y = glu_tail(blob, D2)
                    for m in range(128):
                        mm = m0 + m
                        tl, j = mm // Pp, mm % Pp
                        t = t0 + tl
                        if tl >= nt or t >= T:
                            continue
                        fo = 2 * j + parity
                        if last:
                            if fo < Fo:
                                out[b, t, fo] = y[m]
                        else:
                            out[b, :, t * 2 * Pp + parity * Pp + j, :] = y[m].reshape(8, 8) if fo < Fo else 0.0
    return out


# ------------------------------------------------------------------ TCM
def emu_tcm(blobs, e5, dil):
    """e5 CP8 split F=4 [B][8][T*4][8] -> decoder input CP8 split F=4.  One "launch" per k."""
    B, _, NP, _ = e5.shape
    T = NP // 4
    # launch 0, phase B only: convert to residual-stream order kk = f*64 + c
    x = np.zeros((B, 32, T, 8))
    for f in range(4):
        pos = np.arange(T) * 4 + (f & 1) * 2 + (f >> 1)
        x[:, f * 8:(f + 1) * 8] = e5[:, :, pos, :]

    def phase_b(blob, xb):                    # x [32][T][8] -> am, ak [8][T][8]
        y = np.einsum("ktj,knj->tn", xb, blob.h["w1"]) + blob.f["b1"]
        am = prelu(y, blob.f["slopes"][0]) * blob.f["sm"] + blob.f["shm"]
        ak = prelu(y, blob.f["slopes"][1]) * blob.f["sk"] + blob.f["shk"]
        return am.reshape(T, 8, 8).transpose(1, 0, 2), ak.reshape(T, 8, 8).transpose(1, 0, 2)

    for b in range(B):
        am, ak = phase_b(blobs[0], x[b])
        for k in range(18):
            blob, d = blobs[k], dil[k]
            xn = np.zeros((32, T, 8))
            for t0 in range(0, T, 128):
                # patch rows: time t0-2d .. t0+127+2d  (zeros outside [0,T))
                R = 128 + 4 * d
                pm, pk = np.zeros((8, R, 8)), np.zeros((8, R, 8))
                lo, hi = max(0, t0 - 2 * d), min(T, t0 + 128 + 2 * d)
                pm[:, lo - (t0 - 2 * d):hi - (t0 - 2 * d)] = am[:, lo:hi]
                pk[:, lo - (t0 - 2 * d):hi - (t0 - 2 * d)] = ak[:, lo:hi]
                Dm, Dk = np.zeros((128, 64)), np.zeros((128, 64))
                for tap in range(5):
                    Dm += gemm_planes(pm, tap * d, 128, blob.h["wm"][tap])
                    Dk += gemm_planes(pk, tap * d, 128, blob.h["wk"][tap])
                g = (Dm + blob.f["bm"]) * sigmoid(Dk + blob.f["bk"])
                a3 = prelu(g, blob.f["slopes"][2]) * blob.f["sc"] + blob.f["shc"]
                A3 = a3.reshape(128, 8, 8).transpose(1, 0, 2)
                D3 = gemm_planes(A3, 0, 128, blob.h["w3"]) + blob.f["b3"]      # [128][256]
                n = min(128, T - t0)
                xn[:, t0:t0 + n] = x[b][:, t0:t0 + n] + D3[:n].reshape(n, 32, 8).transpose(1, 0, 2)
            x[b] = xn
            if k < 17:
                am, ak = phase_b(blobs[k + 1], x[b])
    out = np.zeros((B, 8, T * 4, 8))
    for f in range(4):
        pos = np.arange(T) * 4 + (f & 1) * 2 + (f >> 1)
        out[:, :, pos, :] = x[:, f * 8:(f + 1) * 8]
    return out

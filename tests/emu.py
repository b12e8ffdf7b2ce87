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
def bias_rows(block):
    """value a bias MMA adds to every row: ones-plane (1,1,0,..) x block[0][n][0:2]"""
    return block[0, :, 0] + block[0, :, 1]


def glu_tail(blob, D2):
    """D2 [M][128] = l | r | lm' | rm' accumulators (without the bias MMA) -> block output [M][64] (or [M] for de1)."""
    D2 = D2 + bias_rows(blob.h["b_lr4"])
    l, r, tl, tr = D2[:, :32], D2[:, 32:64], np.tanh(D2[:, 64:96]), np.tanh(D2[:, 96:128])
    g = l * tr + l + (r * tl + r)                                 # = 2 * (l * sigmoid_r + r * sigmoid_l)
    if "w2" in blob.h:
        A3 = g.reshape(-1, 4, 8).transpose(1, 0, 2)
        D4 = gemm_planes(A3, 0, A3.shape[1], blob.h["w2"]) + bias_rows(blob.h["b_out"])
        return prelu(D4, blob.f["slope"][0])
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
                D2 = np.zeros((128, 128))
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
                    D2 = np.zeros((128, 128))
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
        y = np.einsum("ktj,knj->tn", xb, blob.h["w1"]) + bias_rows(blob.h["b_1"])
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
                Dm, Dk = Dm + bias_rows(blob.h["b_m"]), Dk + bias_rows(blob.h["b_k"])
                g = Dm * np.tanh(Dk) + Dm                        # = 2 * main * sigmoid(mask); the 1/2 sits in sc
                a3 = prelu(g, blob.f["slopes"][2]) * blob.f["sc"] + blob.f["shc"]
                A3 = a3.reshape(128, 8, 8).transpose(1, 0, 2)
                D3 = gemm_planes(A3, 0, 128, blob.h["w3"]) + bias_rows(blob.h["b_3"])      # [128][256]
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


# =============================================================================
# GCRN prior: layouts + streaming-GEMM emulation (mirrors csrc/gcrn.cu)
# =============================================================================
def elu(x):
    return np.where(x > 0, x, np.exp(np.minimum(x, 0)) - 1)


def to_so(x):
    """[B,C,T,F] -> split-outer CP8 [B][C/8][2][T*Q][8]; row(t, f) = t*Q + (f>>1) in parity plane f&1."""
    B, C, T, F = x.shape
    Q = (F + 1) // 2
    out = np.zeros((B, C // 8, 2, T * Q, 8))
    for f in range(F):
        out[:, :, f & 1, np.arange(T) * Q + (f >> 1), :] = x[:, :, :, f].reshape(B, C // 8, 8, T).transpose(0, 1, 3, 2)
    return out


def to_ug(x):
    """[B,C,T,F] -> unsplit guarded CP8 [B][C/8][T*P+1][8], P = F+1; row(t, f) = t*P + 1 + f."""
    B, C, T, F = x.shape
    Pp = F + 1
    out = np.zeros((B, C // 8, T * Pp + 1, 8))
    for f in range(F):
        out[:, :, np.arange(T) * Pp + 1 + f, :] = x[:, :, :, f].reshape(B, C // 8, 8, T).transpose(0, 1, 3, 2)
    return out


def stream_gemm(A, taps, wstream, ntile, nrows):
    """A [NC][npar][rows][8]; taps [(par, shift)]; wstream: one n-tile's [tap][NC][ntile][8] flat."""
    NC = A.shape[0]
    W = wstream.reshape(len(taps), NC, ntile, 8)
    D = np.zeros((nrows, ntile))
    for ti, (par, sh) in enumerate(taps):
        Aw = np.zeros((NC, nrows, 8))
        avail = max(0, min(nrows, A.shape[2] - sh))
        Aw[:, :avail] = A[:, par, sh:sh + avail]
        D += np.einsum("kmj,knj->mn", Aw, W[ti])
    return D


def glu_bn_elu(D, ep, ntile):
    ct = ntile // 2
    bv, bg, s, sh = ep[0:ct], ep[ct:2 * ct], ep[2 * ct:3 * ct], ep[3 * ct:4 * ct]
    y = (D[:, :ct] + bv) * sigmoid(D[:, ct:] + bg)
    return elu(y * s + sh)


def emu_gcrn_enc(blob, xin_so, i):
    """conv{i} (i>=2) on split-outer input -> dense [B,Cout,T,Fo]."""
    cin, cout = P.GCRN_CH[i - 1], P.GCRN_CH[i]
    Fin, Fo = P.GCRN_F[i - 1], P.GCRN_F[i]
    Q = (Fin + 1) // 2
    B = xin_so.shape[0]
    T = xin_so.shape[3] // Q
    ntile = min(256, 2 * cout)
    ct = ntile // 2
    taps = [(0, 0), (1, 0), (0, 1)]
    out = np.zeros((B, cout, T, Fo))
    wsz = 3 * (cin // 8) * ntile * 8
    for b in range(B):
        for j in range(2 * cout // ntile):
            D = stream_gemm(xin_so[b], taps, blob.h["w"][j * wsz:(j + 1) * wsz], ntile, T * Q)
            y = glu_bn_elu(D, blob.f["ep"][j * 4 * ct:(j + 1) * 4 * ct], ntile)     # [T*Q][ct]
            out[b, j * ct:(j + 1) * ct] = y.reshape(T, Q, ct)[:, :Fo].transpose(2, 0, 1)
    return out


def emu_gcrn_conv1(blob, y_in):
    B, _, T, F = y_in.shape
    out = np.zeros((B, 16, T, 80))
    W = blob.h["w"]                                   # [2][32][8]
    for b in range(B):
        A = np.zeros((T * 80, 16))
        for c in range(2):
            for df in range(3):
                A[:, c * 3 + df] = y_in[b, c][:, df:df + 159:2].reshape(-1)
        D = np.einsum("mkj,knj->mn", A.reshape(-1, 2, 8), W)
        out[b] = glu_bn_elu(D, blob.f["ep"], 32).reshape(T, 80, 16).transpose(2, 0, 1)
    return out


def emu_gcrn_dec(blob, a_ug, b_ug, i):
    """conv{i}_t on [prev | skip] (both unsplit-guarded) -> dense [B,Cout,T,Fout] (BN+ELU applied)."""
    cin, cout, Fin, Fout = P.GCRN_DEC[i]
    Pp = Fin + 1
    B = a_ug.shape[0]
    T = (a_ug.shape[2] - 1) // Pp
    ntile = 2 * cout
    out = np.zeros((B, cout, T, Fout))
    for b in range(B):
        A = np.concatenate([a_ug[b], b_ug[b]], axis=0)[:, None]          # [NC][1][rows][8]
        De = stream_gemm(A, [(0, 1), (0, 0)], blob.h["w_even"], ntile, T * Pp)
        Do = stream_gemm(A, [(0, 1)], blob.h["w_odd"], ntile, T * Pp)
        ye = glu_bn_elu(De, blob.f["ep"], ntile).reshape(T, Pp, cout)
        yo = glu_bn_elu(Do, blob.f["ep"], ntile).reshape(T, Pp, cout)
        for j in range(Pp):
            if 2 * j < Fout:
                out[b, :, :, 2 * j] = ye[:, j].T
            if 2 * j + 1 < Fout:
                out[b, :, :, 2 * j + 1] = yo[:, j].T
    return out


def emu_lstm(blob, x_rows, B, T):
    """x_rows [T*B][512] in the kernel's K order (row = t*B + b) -> h [T*B][512] (unit order)."""
    W = blob.h["w_ih"].reshape(16, 64, 128, 8)                            # [ntile][kc][128][8]
    A = x_rows.reshape(T * B, 64, 8).transpose(1, 0, 2)
    pre = np.concatenate([np.einsum("kmj,knj->mn", A, W[j]) for j in range(16)], axis=1) + blob.f["bias"]
    pre = pre.reshape(T, B, 2048)
    Whh = blob.h["w_hh"]                                                  # [16][64][128][8]
    h = np.zeros((B, 512))
    c = np.zeros((B, 512))
    out = np.zeros((T, B, 512))
    for t in range(T):
        hp = h.reshape(B, 64, 8).transpose(1, 0, 2)                       # B operand [kc][b][8]
        hn = np.zeros_like(h)
        for cta in range(16):
            D = np.einsum("kmj,knj->mn", Whh[cta], hp) + pre[t][:, cta * 128:(cta + 1) * 128].T   # [128 lanes][B]
            gi, gf, gg, go = sigmoid(D[0:32]), sigmoid(D[32:64]), np.tanh(D[64:96]), sigmoid(D[96:128])
            u = slice(cta * 32, cta * 32 + 32)
            c[:, u] = (gf * c[:, u].T + gi * gg).T
            hn[:, u] = (go * np.tanh(c[:, u].T)).T
        h = hn
        out[t] = h
    return out.reshape(T * B, 512)


def layer_norm(x, w, b):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + 1e-5) * w + b


def emu_gcrn(pk, y_in):
    """whole GCRN forward through the kernel layouts -> X_init (= out/11) [B,2,T,161]."""
    B, _, T, _ = y_in.shape
    e = [None, emu_gcrn_conv1(pk["conv1"], y_in)]
    for i in range(2, 6):
        e.append(emu_gcrn_enc(pk[f"conv{i}"], to_so(e[i - 1]), i))
    # LSTM layer 1 input: row = t*B + b, kk = f*128 + cl
    e5 = e[5]                                                             # [B,256,T,4]
    hs = []
    for g in range(2):
        xg = e5[:, 128 * g:128 * (g + 1)].transpose(2, 0, 3, 1).reshape(T * B, 512)   # [t][b][f][cl]
        hs.append(emu_lstm(pk[f"lstm1_{g}"], xg, B, T))
    inter = np.stack(hs, axis=-1).reshape(T * B, 1024)                    # feature' = 2j + g
    ln1 = layer_norm(inter, pk["ln"].f["w1"], pk["ln"].f["b1"])
    hs = [emu_lstm(pk[f"lstm2_{g}"], ln1[:, 512 * g:512 * (g + 1)], B, T) for g in range(2)]
    ln2 = layer_norm(np.concatenate(hs, axis=-1), pk["ln"].f["w2"], pk["ln"].f["b2"])   # [T*B][1024], c*4+f
    lstm_out = ln2.reshape(T, B, 256, 4).transpose(1, 2, 0, 3)            # [B,256,T,4]
    outs = []
    for br in (1, 2):
        d = emu_gcrn_dec(pk[f"dec{br}_5"], to_ug(lstm_out), to_ug(e5), 5)
        for i in range(4, 1, -1):
            d = emu_gcrn_dec(pk[f"dec{br}_{i}"], to_ug(d), to_ug(elu(e[i])), i)
        # conv1_t (32 -> 1 GLU) + bn + elu + fc, on CUDA cores in the kernel
        ob = pk[f"out{br}"].f
        cat = np.concatenate([d, elu(e[1])], axis=1)                      # [B,32,T,80]
        wv, wg = ob["wv"].reshape(32, 3), ob["wg"].reshape(32, 3)
        bv, bg, s, sh = ob["misc"]
        d1 = np.zeros((B, T, 161))
        for nm, w, bias in (("v", wv, bv), ("g", wg, bg)):
            acc = np.full((B, T, 161), bias)
            for j in range(80):
                acc[:, :, 2 * j] += np.einsum("bct,c->bt", cat[:, :, :, j], w[:, 0])
                acc[:, :, 2 * j + 1] += np.einsum("bct,c->bt", cat[:, :, :, j], w[:, 1])
                acc[:, :, 2 * j + 2] += np.einsum("bct,c->bt", cat[:, :, :, j], w[:, 2])
            if nm == "v":
                val = acc
            else:
                d1 = elu(val * sigmoid(acc) * s + sh)
        outs.append(d1 @ ob["fcw"].reshape(161, 161) + ob["fcb"][:161])
    return np.stack(outs, axis=1)


# ============================================================================ warp-level FFT of csrc/signal.cu
# The STFT / ISTFT kernels transform one frame per warp: 320 real samples = 160 complex points = 5 x 32; lane j holds
# z[j + 32 m2], radix-5 in registers, W160^(j q) twiddles, five 32-point FFTs across the lanes by butterfly shuffles
# (DIF: natural in, bit-reversed out; the inverse is DIT).  These functions re-run that lane arithmetic in float32 NumPy
# (arrays indexed by lane, shfl_xor = index permutation) with the same table the host builds.
_LANE = np.arange(32)
_BREV = np.array([int("{:05b}".format(l)[::-1], 2) for l in range(32)])
_C1, _S1 = np.float32(0.30901699437494745), np.float32(0.9510565162951535)
_C2, _S2 = np.float32(-0.8090169943749473), np.float32(0.5877852522924731)


def fft_tables():
    """Hann[320] and W320^r = exp(-2 pi i r / 320), evaluated in float64 and rounded to float32 (pdse_signal_tables)"""
    n = np.arange(320)
    hann = (0.5 - 0.5 * np.cos(2 * np.pi * n / 320)).astype(np.float32)
    tw = (np.cos(2 * np.pi * n / 320).astype(np.float32) - 1j * np.sin(2 * np.pi * n / 320).astype(np.float32)).astype(np.complex64)
    return hann, tw


def _radix5(x, inverse):
    x0, x1, x2, x3, x4 = x
    t1, t2, t3, t4 = x1 + x4, x2 + x3, x1 - x4, x2 - x3
    a1 = x0 + _C1 * t1 + _C2 * t2
    a2 = x0 + _C2 * t1 + _C1 * t2
    b1 = _S1 * t3 + _S2 * t4
    b2 = _S2 * t3 - _S1 * t4
    y0 = x0 + t1 + t2
    ib1, ib2 = (1j * b1).astype(np.complex64), (1j * b2).astype(np.complex64)
    if inverse:
        return [y0, a1 + ib1, a2 + ib2, a2 - ib2, a1 - ib1]
    return [y0, a1 - ib1, a2 - ib2, a2 + ib2, a1 + ib1]


def _stage_tw(tw, h):
    return np.where((_LANE & h) != 0, tw[(10 * (16 // h) * (_LANE & (h - 1))) % 320], np.complex64(1))


def _fft32_dif(v, tw):
    for h in (16, 8, 4, 2, 1):
        p = v[_LANE ^ h]
        v = np.where((_LANE & h) != 0, (p - v) * _stage_tw(tw, h), v + p).astype(np.complex64)
    return v


def _ifft32_dit(v, tw):
    for h in (1, 2, 4, 8, 16):
        bit = (_LANE & h) != 0
        y = np.where(bit, v * np.conj(_stage_tw(tw, h)), v).astype(np.complex64)
        p = y[_LANE ^ h]
        v = np.where(bit, p - y, y + p).astype(np.complex64)
    return v


def stft_frame_lanes(samples):
    """320 un-windowed float32 samples -> X[0..160] (complex64), as stft_compress_kernel computes one frame"""
    hann, tw = fft_tables()
    xw = samples.astype(np.float32) * hann
    z = (xw[0::2] + 1j * xw[1::2]).astype(np.complex64)
    y = _radix5([z[_LANE + 32 * m] for m in range(5)], False)
    Z = np.zeros(160, np.complex64)
    for q in range(5):
        v = y[q] * tw[(2 * _LANE * q) % 320] if q else y[q]
        Z[5 * _BREV + q] = _fft32_dif(v.astype(np.complex64), tw)
    X = np.zeros(161, np.complex64)
    for k in range(81):
        A, B = Z[k], np.conj(Z[(160 - k) % 160])
        E, O = np.complex64(0.5) * (A + B), np.complex64(-0.5j) * (A - B)
        X[k] = E + tw[k] * O
        if k != 80:
            X[160 - k] = np.conj(E - tw[k] * O)
    return X


def istft_frame_lanes(X):
    """X[0..160] -> 320 float32 samples windowed by the synthesis Hann (irfft semantics), as decompress_istft_kernel"""
    hann, tw = fft_tables()
    X = X.astype(np.complex64).copy()
    X[0], X[160] = X[0].real, X[160].real
    Z = np.zeros(160, np.complex64)
    for k in range(81):
        A, B = X[k], np.conj(X[160 - k])
        E, O = np.complex64(0.5) * (A + B), np.conj(tw[k]) * (np.complex64(0.5) * (A - B))
        Z[k] = E + 1j * O
        if 0 < k < 80:
            Z[160 - k] = np.conj(E) + 1j * np.conj(O)
    y = []
    for q in range(5):
        v = _ifft32_dit(Z[5 * _BREV + q], tw)
        y.append((v * np.conj(tw[(2 * _LANE * q) % 320])).astype(np.complex64) if q else v)
    z5 = _radix5(y, True)
    x = np.zeros(320, np.float32)
    for m in range(5):
        x[2 * (_LANE + 32 * m)] = z5[m].real * (hann[2 * (_LANE + 32 * m)] * np.float32(1 / 160))
        x[2 * (_LANE + 32 * m) + 1] = z5[m].imag * (hann[2 * (_LANE + 32 * m) + 1] * np.float32(1 / 160))
    return x


# ---------------------------------------------------------------------------------------------- diff2.DiffWave (csrc/diffwave.cu)
def emu_diffwave(pk, win, wout_post, audio, init, e_rows):
    """NumPy statement of what the DiffWave kernels compute from the PACKED operands (prior_diffuse_b200.diffwave.pack_diffwave):
    the K order of W_cat (phase, tap, channel), the chunk-plane layouts, the bias blocks, the zero guard rows and the
    skip / output projections.  audio, init [L]; e_rows [layers][64] = the per-layer diffusion projections.  -> out [L]"""
    L = audio.shape[0]
    G = 640                                                   # guard rows (pdse_dw_guard_rows)
    w_in, b_in = win[:64], win[64:]
    x = np.maximum(audio[:, None] * w_in[None, :] + b_in[None, :], 0.0)            # [L][64]
    cond = np.zeros((L + 2 * G, 64))
    cond[G:G + L] = np.maximum(init[:, None] * w_in[None, :] + b_in[None, :], 0.0)
    skip = np.zeros((L, 64))
    for i, blob in enumerate(pk["blobs"]):
        d = pk["dilations"][i]
        wcat = blob[:48 * 128 * 8].reshape(48, 128, 8)        # [K/8][N][8]
        wo = blob[48 * 128 * 8:56 * 128 * 8].reshape(8, 128, 8)
        b_conv = blob[56 * 128 * 8:56 * 128 * 8 + 2 * 128 * 8].reshape(2, 128, 8)
        b_out = blob[56 * 128 * 8 + 2 * 128 * 8:].reshape(2, 128, 8)
        Wcat = wcat.transpose(1, 0, 2).reshape(128, 384)      # [N][K]
        Wo = wo.transpose(1, 0, 2).reshape(128, 64)
        y = np.zeros((L + 2 * G, 64))
        y[G:G + L] = x + e_rows[i][None, :]
        # six windows: operand (y, cond) x tap (t - d, t, t + d); K = (phase * 3 + tap) * 64 + c
        A = np.concatenate([src[G + off:G + off + L] for src in (y, cond) for off in (-d, 0, d)], axis=1)      # [L][384]
        z = A @ Wcat.T + (b_conv[0, :, 0] + b_conv[0, :, 1])[None, :]
        g = sigmoid(z[:, :64]) * np.tanh(z[:, 64:])
        o = g @ Wo.T + (b_out[0, :, 0] + b_out[0, :, 1])[None, :]
        x = (x + o[:, :64]) / np.sqrt(2.0)
        skip = skip + o[:, 64:]
    post = pk["post"]
    Wk = post[:8 * 64 * 8].reshape(8, 64, 8).transpose(1, 0, 2).reshape(64, 64)
    bk = post[8 * 64 * 8:].reshape(2, 64, 8)
    h = np.maximum((skip / np.sqrt(len(pk["blobs"]))) @ Wk.T + (bk[0, :, 0] + bk[0, :, 1])[None, :], 0.0)
    return h @ wout_post[:64] + wout_post[64]

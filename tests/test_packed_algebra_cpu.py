"""CPU model of the time-packed whole-ResBlock kernel (lip2speech-unit_b200/csrc/respk_tc.cuh): block-Toeplitz weight
expansion (pk_pack_weights), the offset MMAs over two 128-byte-row slabs with row shifts, phase-major block order for the
dilated convs, halo accounting (mt_v, h_tot, hl, r_out), the c2 biases kept out of the accumulator and the tile walk --
restated in numpy (float64) and checked against a direct ResBlock1 (speech-resynthesis/models.py:34-41).  It pins the
ALGEBRA the CUDA kernel implements; the kernel itself is checked on the GPU (tests/test_gpu_forward.py)."""
import numpy as np
import pytest

rng = np.random.default_rng(0)
PAD = 8
def pack_weights(w, c, k):
    P = 128 // c; Q = P // 2
    n_off = P + k - 1; n_groups = (n_off + Q - 1) // Q
    out = np.zeros((n_groups, 128, 64), np.float64)
    for oi in range(n_off):
        gq, s = divmod(oi, Q)
        for j in range(P):
            tap = oi - j
            if tap < 0 or tap >= k: continue
            out[gq, j*c:(j+1)*c, s*c:(s+1)*c] = w[:, :, tap]   # [co][ci]
    return out
def pos(d, span, tau):
    if d == 1: return tau
    p, r = divmod(tau, d)
    return r*span + p if p < span else -1
def tau_of(d, span, x):
    if d == 1: return x
    r, p = divmod(x, span)
    return p*d + r if r < d else -1
def conv_packed(S, wt, c, k, nr):
    """S: dict half -> array [(nr+2*PAD)][64] ; returns D [nr][128]"""
    P = 128 // c; Q = P // 2; hc = (k-1)//2
    n_off = P + k - 1
    D = np.zeros((nr, 128))
    for oi in range(n_off):
        gq, s = divmod(oi, Q)
        o = oi - hc
        shift = o // P; e = o % P
        h, sl = divmod(e, Q)
        A = S[h][PAD+shift:PAD+shift+nr, sl*c:(sl+1)*c]          # [nr][c]
        B = wt[gq][:, s*c:(s+1)*c]                                # [128][c]
        D += A @ B.T
    return D
def store(S, vals, c, src_nat, d, span, nr, t_row0, lin, edge=True):
    """vals [nr][128] -> lrelu -> S in dst layout"""
    P = 128 // c; Q = P // 2
    v = np.where(vals > 0, vals, 0.1*vals)
    for row in range(nr):
        for j in range(P):
            xs = row*P + j
            if src_nat: tau = xs; xd = pos(d, span, tau)
            else: tau = tau_of(d, span, xs); xd = tau
            if tau < 0 or xd < 0: continue
            piece = v[row, j*c:(j+1)*c]
            t = t_row0 + tau
            if edge and (t < 0 or t >= lin): piece = np.zeros(c)
            drow, de = divmod(xd, P); h, sl = divmod(de, Q)
            S[h][PAD+drow, sl*c:(sl+1)*c] = piece
def ref_resblock(x, W1, B1, W2, B2, dil, k):
    # x [L][c]
    def conv(a, w, b, d):
        L, c = a.shape; h = (k-1)//2
        ap = np.zeros((L + 2*h*d, c)); ap[h*d:h*d+L] = a
        out = np.zeros((L, c))
        for tap in range(k):
            out += ap[tap*d:tap*d+L] @ w[:, :, tap].T
        return out + b
    lr = lambda a: np.where(a > 0, a, 0.1*a)
    for s, d in enumerate(dil):
        t = conv(lr(x), W1[s], B1[s], d)
        t = conv(lr(t), W2[s], B2[s], 1)
        x = x + t
    return x
def run(c, k, dil, msub, lin):
    P = 128 // c; nr = 128*msub; mt = nr*P; hc = (k-1)//2
    lays = []; mt_v = mt; h_tot = 0
    for d in dil:
        nb = nr // d
        lays.append((d, mt if d == 1 else nb*P))
        mt_v = min(mt_v, d*nb*P); h_tot += (d+1)*hc
    hl = (h_tot + P - 1)//P*P
    r_out = (mt_v - h_tot - hl)//P*P
    m_items = (lin + r_out - 1)//r_out
    W1 = [rng.normal(size=(c, c, k))*0.2 for _ in dil]; W2 = [rng.normal(size=(c, c, k))*0.2 for _ in dil]
    B1 = [rng.normal(size=c)*0.1 for _ in dil]; B2 = [rng.normal(size=c)*0.1 for _ in dil]
    x = rng.normal(size=(lin, c))
    want = ref_resblock(x, W1, B1, W2, B2, dil, k)
    got = np.full((lin, c), np.nan)
    wt1 = [pack_weights(w, c, k) for w in W1]; wt2 = [pack_weights(w, c, k) for w in W2]
    for mi in range(m_items):
        q0 = mi*r_out; t_row0 = q0 - hl
        S = {0: rng.normal(size=(nr+2*PAD, 64))*0 , 1: np.zeros((nr+2*PAD, 64))}
        X = np.zeros((nr, 128))
        for row in range(nr):
            t = t_row0 + row*P
            if 0 <= t < lin: X[row] = x[t:t+P].reshape(-1)
        store(S, X, c, True, lays[0][0], lays[0][1], nr, t_row0, lin, edge=False)
        run_b2 = np.zeros(c)
        for s, d in enumerate(dil):
            D1 = conv_packed(S, wt1[s], c, k, nr)
            store(S, D1 + np.tile(B1[s], P), c, lays[s][0] == 1, lays[s][0] if lays[s][0] != 1 else 1, lays[s][1] if lays[s][0] != 1 else mt, nr, t_row0, lin)
            X = X + conv_packed(S, wt2[s], c, k, nr)
            run_b2 = run_b2 + B2[s]
            if s + 1 < len(dil):
                store(S, X + np.tile(run_b2, P), c, True, lays[s+1][0], lays[s+1][1], nr, t_row0, lin)
        out = (X + np.tile(run_b2, P)).reshape(mt, c)
        lo, hi = q0, min(lin, q0 + r_out)
        got[lo:hi] = out[lo - t_row0:hi - t_row0]
    err = np.abs(got - want).max()
    # print(f"c={c} k={k} dil={dil} msub={msub} lin={lin}: mt={mt} mt_v={mt_v} h_tot={h_tot} hl={hl} r_out={r_out} items={m_items} max err {err:.2e}")
    assert err < 1e-9


@pytest.mark.parametrize("c,k,msub,lin", [(16, 3, 1, 1280), (16, 11, 1, 1280), (32, 7, 1, 1280), (64, 11, 1, 384), (64, 3, 2, 700),
                                          (16, 11, 2, 4992), (32, 7, 1, 8)])
def test_packed_resblock_algebra(c, k, msub, lin):
    run(c, k, [1, 3, 5], msub, lin)

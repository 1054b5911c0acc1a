"""Thread-accurate Python emulation of csrc/seqsum2.cuh::block_seqsum_exact_v2_t (T threads, warp shuffles with CUDA semantics)
against the literal float32 loop, on the exact inputs of tests/test_gpu_parity.py::test_exact_parallel_sequential_sum.
Usage: python tools/seqsum2/emulate_v2.py [T]"""
import sys

import numpy as np

F = np.float32
LITERAL = -(2 ** 31)


def f32_exponent(f):
    return int((np.float32(f).view(np.uint32) >> 23) & 0xFF) - 127


def seq_pair(t, e):
    tb = int(np.float32(t).view(np.uint32))
    et = tb >> 23
    if et == 0:
        return True, (0, 0)
    m = (tb & 0x7FFFFF) | 0x800000
    shift = (e + 127) - et
    if shift < 0:
        return False, (0, 0)
    if shift > 25:
        shift = 25
    k = m >> shift
    rem = m & ((1 << shift) - 1)
    half = (1 << (shift - 1)) if shift else 0
    if shift and rem == half:
        return True, (k + (k & 1), k + ((k + 1) & 1))
    v = k + (1 if (shift and rem > half) else 0)
    return True, (v, v)


def compose(L, R):
    a0 = L[0] + (R[1] if (L[0] & 1) else R[0])
    a1 = L[1] + (R[1] if ((1 + L[1]) & 1) else R[0])
    return (min(a0, 1 << 26), min(a1, 1 << 26))


def warp_segscan(ps, fs):
    """ps, fs: lists of 32 (pair, flag); CUDA __shfl_up semantics (lane < d keeps its own value)."""
    ps, fs = list(ps), list(fs)
    d = 1
    while d < 32:
        ups = [ps[l - d] if l >= d else ps[l] for l in range(32)]
        ufs = [fs[l - d] if l >= d else fs[l] for l in range(32)]
        for l in range(32):
            if l >= d and not fs[l]:
                ps[l] = compose(ups[l], ps[l])
                fs[l] = ufs[l]
        d <<= 1
    return ps, fs


def v2(terms, T):
    n = len(terms)
    E = (n + T - 1) // T
    sq = np.zeros(T * E, dtype=F)
    sq[:n] = terms
    NW = T // 32
    loc = np.zeros(T, dtype=F)
    for t in range(T):
        s = F(0)
        for k in range(E):
            s = F(s + sq[t * E + k])
        loc[t] = s
    inc = loc.copy()
    for w in range(NW):
        v = inc[w * 32:(w + 1) * 32].copy()
        d = 1
        while d < 32:
            u = np.concatenate([v[:d], v[:-d]])
            for l in range(32):
                if l >= d:
                    v[l] = F(v[l] + u[l])
            d <<= 1
        inc[w * 32:(w + 1) * 32] = v
    wsum = np.zeros(32, dtype=F)
    for w in range(NW):
        wsum[w] = inc[w * 32 + 31]
    v = wsum.copy()
    d = 1
    while d < 32:
        u = np.concatenate([v[:d], v[:-d]])
        for l in range(32):
            if l >= d:
                v[l] = F(v[l] + u[l])
        d <<= 1
    wex = np.array([F(v[l] - wsum[l]) for l in range(32)], dtype=F)
    cls = [LITERAL] * T
    pr = [(0, 0)] * T
    for t in range(T):
        w = t >> 5
        p_end = F(wex[w] + inc[t])
        p_start = F(wex[w] + F(inc[t] - loc[t]))
        e = f32_exponent(p_start)
        if p_start > 0 and e > -100 and e < 128 and f32_exponent(p_end) == e:
            b = np.uint32((e + 127) << 23).view(F) if 0 <= e + 127 < 256 else F(np.inf)
            with np.errstate(all="ignore"):
                lo = F(b * F(1.0 + 2.0 ** -9))
                hi = F(F(2.0) * b) * F(1.0 - 2.0 ** -9)
            if p_start >= lo and p_end <= F(hi):
                ok, p = True, (0, 0)
                for k in range(E):
                    o, q = seq_pair(sq[t * E + k], e)
                    if not o:
                        ok = False
                        break
                    p = compose(p, q)
                if ok:
                    cls[t] = e
                    pr[t] = p
    f = [0] * T
    isitem = [False] * T
    for t in range(T):
        prev = cls[t - 1] if t > 0 else LITERAL
        nxt = cls[t + 1] if t < T - 1 else LITERAL
        clean = cls[t] != LITERAL
        f[t] = 0 if (clean and prev == cls[t]) else 1
        isitem[t] = (not clean) or nxt != cls[t]
    wtail, wtail_f = [None] * NW, [None] * NW
    for w in range(NW):
        ps, fs = warp_segscan(pr[w * 32:(w + 1) * 32], f[w * 32:(w + 1) * 32])
        pr[w * 32:(w + 1) * 32] = ps
        f[w * 32:(w + 1) * 32] = fs
        wtail[w], wtail_f[w] = ps[31], fs[31]
    tp = [(wtail[l] if l < NW else (0, 0)) for l in range(32)]
    tf = [(wtail_f[l] if l < NW else 1) for l in range(32)]
    tp, tf = warp_segscan(tp, tf)
    for t in range(T):
        w = t >> 5
        if not f[t] and w > 0:
            pr[t] = compose(tp[w - 1], pr[t])
    items = [(cls[t], pr[t], t) for t in range(T) if isitem[t]]
    s = F(0)
    fallbacks = 0
    with np.errstate(all="ignore"):
        for c, p, last in items:
            if c == LITERAL:
                for k in range(E):
                    s = F(s + sq[last * E + k])
                continue
            sb = int(np.float32(s).view(np.uint32))
            ok = f32_exponent(s) == c and (sb >> 23) != 0
            if ok:
                M = (sb & 0x7FFFFF) | 0x800000
                M2 = M + (p[1] if (M & 1) else p[0])
                if M2 < (1 << 24):
                    s = np.uint32((((c + 127) << 23) | (M2 & 0x7FFFFF)) & 0xFFFFFFFF).view(F)
                else:
                    ok = False
            if not ok:
                first = last
                while first > 0 and cls[first - 1] == c:
                    first -= 1
                for k in range(first * E, (last + 1) * E):
                    s = F(s + sq[k])
                fallbacks += 1
    return s, len(items), fallbacks


def cases():
    rng = np.random.default_rng(0)
    out = []
    for trial in range(120):
        n = int(rng.choice([33, 64, 256, 1000, 1536, 2560, 4096, 8192]))
        kind = trial % 10
        if kind == 0: x = rng.standard_normal(n)
        elif kind == 1: x = rng.standard_normal(n) * 10.0 ** rng.integers(-6, 6)
        elif kind == 2: x = rng.standard_cauchy(n)
        elif kind == 3: x = np.full(n, rng.standard_normal())
        elif kind == 4: x = 2.0 ** rng.integers(-10, 10, n)
        elif kind == 5:
            x = rng.standard_normal(n); x[rng.integers(0, n, n // 4)] = 0
        elif kind == 6: x = rng.standard_normal(n) * np.exp(rng.standard_normal(n) * 3)
        elif kind == 7: x = np.round(rng.standard_normal(n) * 8) / 8
        elif kind == 8:
            x = np.zeros(n); x[n // 2:] = rng.standard_normal(n - n // 2)
        else:
            x = rng.standard_normal(n) * 1e-3; x[rng.integers(min(40, n - 1), n)] = 1e3
        out.append((x.astype(np.float32) ** 2).astype(np.float32))
    for n in (512, 4096):
        for eps in (-3e-7, -1e-7, 0.0, 1e-7, 3e-7):
            t = np.full(n, (1.0 + eps) / n, dtype=np.float64).astype(np.float32)
            out.append(t)
            out.append(np.concatenate([t, t]).astype(np.float32)[: min(2 * n, 8192)])
    out.append((4.0 ** (np.arange(300) % 150 - 75)).astype(np.float32))
    out.append(np.array([1.0] * 40 + [np.inf] + [1.0] * 40, dtype=np.float32))
    return out


if __name__ == "__main__":
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    bad = 0
    for i, t in enumerate(cases()):
        with np.errstate(all="ignore"):
            ref = np.add.accumulate(t, dtype=np.float32)[-1]
        got, items, fb = v2(t, T)
        same = np.float32(got).view(np.uint32) == np.float32(ref).view(np.uint32) or (np.isnan(got) and np.isnan(ref))
        if not same:
            bad += 1
            print(f"case {i}: n={len(t)} got={got!r} ref={ref!r} items={items} fallbacks={fb}")
    print(f"T={T}: {bad} mismatches")

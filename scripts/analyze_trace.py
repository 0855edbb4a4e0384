import numpy as np, sys, collections
d = np.load('gpurun_out/trace_pair.npz'); ev, n = d['ev'], d['n']
def stream(s):
    e = ev[s, :n[s]]
    return list(zip((e & 255).tolist(), (e >> 8).tolist()))
import os
NTL = int(os.environ.get("NTL", 2))
N, L, T = 10, 3, 10
def tmin(m): return max(0, 9 - m)
# update warp replay
def replay_upd(s):
    it = iter(stream(s)); out = []
    def take(k):
        kk, c = next(it); assert kk == k, (kk, k, len(out)); return c
    for m in range(N):
        for l in range(L):
            for X in range(NTL): a = take(18); b = take(14); out.append(('fpro', m, l, X, 0, dict(s=a, e=b)))
            for t in range(T):
                for X in range(NTL):
                    r = dict(s=take(16), sw=take(11), w0=take(17), w1=take(1), pw=take(24), st=take(25))
                    if t + 1 < T: r['u'] = take(2); r['e'] = take(3)
                    else: r['e'] = take(13); r['u'] = r['e']
                    out.append(('f', m, l, X, t, r))
        take(19); take(15)
    for m in range(N - 1, -1, -1):
        take(21); take(15)
        for l in range(L - 1, -1, -1):
            for t in range(T - 1, tmin(m) - 1, -1):
                for X in range(NTL):
                    r = dict(s=take(20), sw=take(11))
                    if t < T - 1: r['w0'] = take(12); r['w1'] = take(6)
                    else: r['w0'] = r['w1'] = r['sw']
                    r['u'] = take(7); r['e'] = take(8)
                    out.append(('b', m, l, X, t, r))
            for X in range(NTL): a = take(22); b = take(6); c = take(14); out.append(('btail', m, l, X, 0, dict(s=a, w1=b, e=c)))
    return out
def replay_iss():
    it = iter(stream(0)); out = []
    def take(k):
        kk, c = next(it); assert kk == k, (kk, k, len(out)); return c
    for m in range(N):
        for l in range(L):
            for X in range(NTL): a = take(18); b = take(14); out.append(('fpro', m, l, X, 0, dict(s=a, e=b)))
            for t in range(T):
                for X in range(NTL):
                    r = dict(s=take(23), f=take(1))
                    if t + 1 < T: r['r'] = take(4); r['i'] = take(5)
                    out.append(('f', m, l, X, t, r))
        take(19); take(15)
    for m in range(N - 1, -1, -1):
        take(21); take(15)
        for l in range(L - 1, -1, -1):
            for t in range(T - 1, tmin(m) - 1, -1):
                for X in range(NTL):
                    r = {}
                    if t < T - 1: r['s'] = take(23); r['f'] = take(6)
                    r['r'] = take(9); r['i'] = take(10)
                    out.append(('b', m, l, X, t, r))
            for X in range(NTL): a = take(22); b = take(6); c = take(14); out.append(('btail', m, l, X, 0, dict(s=a, w1=b, e=c)))
    return out
U5, U13, I = replay_upd(1), replay_upd(2), replay_iss()
t0 = I[0][5]['s']
print("total cycles", U13[-1][5]['e'] - t0)
# forward: waits by t
for name, U in (('w5', U5), ('w13', U13)):
    for kind in ('f', 'b'):
        acc = collections.defaultdict(list); upd = collections.defaultdict(list); tot = collections.defaultdict(list)
        for (k, m, l, X, t, r) in U:
            if k != kind: continue
            acc[t].append(r['w1'] - r['w0']); upd[t].append(r['u'] - r['w1']); tot[t].append(r['e'] - r['s'])
        print(name, kind, 'wait by t :', {t: int(np.mean(v)) for t, v in sorted(acc.items())})
        print(name, kind, 'upd  by t :', {t: int(np.mean(v)) for t, v in sorted(upd.items())})
        print(name, kind, 'item by t :', {t: int(np.mean(v)) for t, v in sorted(tot.items())})
# print a detailed timeline of one layer: m=5,l=1 forward
def show(kind, m, l):
    print(f"--- timeline {kind} m={m} l={l} (cycles rel. to first event)")
    rows = []
    for nm, U in (('I', I), ('w5', U5), ('w13', U13)):
        for (k, mm, ll, X, t, r) in U:
            if mm == m and ll == l and k.startswith(kind[0]):
                rows.append((min(r.values()), nm, k, X, t, r))
    rows.sort(key=lambda x: x[0]); base = rows[0][0]
    for c, nm, k, X, t, r in rows:
        print(f"{c - base:8d} {nm:4s} {k:5s} X={X} t={t} " + " ".join(f"{a}={b - base}" for a, b in r.items()))
show('f', 5, 1)
show('b', 5, 1)

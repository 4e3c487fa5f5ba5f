#!/usr/bin/env python
"""Distinct 32-byte sectors and 128-byte lines a beam touches, for candidate dose-volume layouts
(VERDICT r1 item 2c), measured on the bench workload with the CPU oracle: uniform(-1,1) actions,
tumour id (i*7919) mod 1000, 100 steps per episode.

A layout is a sector shape (s0, s1, s2) with s0*s1*s2 = 8 voxels (32 B) and a line shape
(l0, l1, l2) in units of sectors with l0*l1*l2 = 4 (128 B).  C order = sector (1,1,8), line (1,1,4).
Reported per layout: sectors / beam, lines / beam, first-touch fraction of sectors and lines over
the episode, distinct sectors / lines per episode, and the DRAM bytes per env-step of a design that
reads and writes every touched sector (RW = 64*Sec) or reads only re-touched ones (D = 32*Sec*(2-first)).
"""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O

G = (67, 43, 70)


def traces(n=64, T=100, kind="uniform", seed=0):
    ph = O.Phantom()
    rng = np.random.default_rng(seed)
    acts = (rng.uniform(-1, 1, (T, n, 6)) if kind == "uniform" else rng.standard_normal((T, n, 6))).astype(np.float32)
    out = []
    for i in range(n):
        env = O.OracleEnv(ph, (i * 7919) % 1000)
        ep = []
        for t in range(T):
            env.step(acts[t, i])
            p = env.pose
            idx, _, _ = O.beam_trace(p[:3], p[3:])
            idx = np.unique(idx)
            c0, r = np.divmod(idx, G[1] * G[2])
            c1, c2 = np.divmod(r, G[2])
            dom = int(np.argmax(np.abs(p[3:].astype(np.float32))))
            ep.append((c0, c1, c2, dom))
        out.append(ep)
    return out


def measure(tr, sec, line):
    """sec = voxels per sector along each axis, line = sectors per line along each axis."""
    ns = [-(-G[a] // sec[a]) for a in range(3)]                   # sectors along each axis
    nl = [-(-ns[a] // line[a]) for a in range(3)]
    SEC, LIN, FS, FL, SEP, LEP = [], [], [], [], [], []
    by_dom = {0: [], 1: [], 2: []}
    for ep in tr:
        seen_s, seen_l = set(), set()
        for c0, c1, c2, dom in ep:
            s0, s1, s2 = c0 // sec[0], c1 // sec[1], c2 // sec[2]
            sid = (s0 * ns[1] + s1) * ns[2] + s2
            lid = ((s0 // line[0]) * nl[1] + s1 // line[1]) * nl[2] + s2 // line[2]
            ss, ls = set(sid.tolist()), set(lid.tolist())
            SEC.append(len(ss)); LIN.append(len(ls))
            by_dom[dom].append(len(ss))
            FS.append(len(ss - seen_s)); FL.append(len(ls - seen_l))
            seen_s |= ss; seen_l |= ls
        SEP.append(len(seen_s)); LEP.append(len(seen_l))
    SEC, LIN, FS, FL = map(lambda x: np.array(x, dtype=np.float64), (SEC, LIN, FS, FL))
    return dict(sec=SEC.mean(), sec_max=SEC.max(), lin=LIN.mean(), lin_max=LIN.max(), first_s=FS.mean() / SEC.mean(),
                first_l=FL.mean() / LIN.mean(), sec_ep=np.mean(SEP), lin_ep=np.mean(LEP),
                dom=[np.mean(by_dom[d]) if by_dom[d] else float("nan") for d in range(3)],
                dom_frac=[len(by_dom[d]) / len(SEC) for d in range(3)],
                vol_bytes=int(np.prod([nl[a] * line[a] * sec[a] for a in range(3)])) * 4)


def shapes(n):
    return [(a, b, n // (a * b)) for a in (1, 2, 4, 8) for b in (1, 2, 4, 8) if a * b <= n and n % (a * b) == 0]


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "uniform"
    tr = traces(kind=kind)
    rows = []
    for sec in shapes(8):
        for line in shapes(4):
            r = measure(tr, sec, line)
            rows.append((sec, line, r))
    rows.sort(key=lambda x: x[2]["sec"] * 1000 + x[2]["lin"])
    print(f"{kind} actions, 64 envs x 100 steps; dominant-axis mix {rows[0][2]['dom_frac']}")
    print("sector   line(sectors)  Sec/beam (max)  by dom axis 0/1/2      Lines/beam (max)  first-touch S/L   per-episode S/L   "
          "RW B/step  D B/step  volume B")
    for sec, line, r in rows:
        rw = 256 + 64 * r["sec"]
        d = 256 + 32 * r["sec"] * (2 - r["first_s"])
        print(f"{sec}  {line}   {r['sec']:6.1f} ({r['sec_max']:3.0f})   {r['dom'][0]:5.1f} {r['dom'][1]:5.1f} {r['dom'][2]:5.1f}   "
              f"{r['lin']:6.1f} ({r['lin_max']:3.0f})   {r['first_s']:.2f} / {r['first_l']:.2f}   {r['sec_ep']:6.0f} / {r['lin_ep']:5.0f}   "
              f"{rw:7.0f}  {d:7.0f}  {r['vol_bytes']}")


if __name__ == "__main__":
    main()

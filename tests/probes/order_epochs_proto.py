"""Prototype (numpy) of the data-parallel form of the reference candidate order that the device code implements
(fastsmc_b200/csrc/order_kernels.cuh): per-word creation order by one global sort, rehash schedule from counts alone,
then one pass per epoch (interval between two rehashes of the reference's extend map): sort by (bucket, time), a
sequential walk per bucket, re-keying of the live nodes at the rehash.  Checked against the literal replay
(CandidateOrder.hpp: replayReferenceOrder).  Run: python tests/probes/order_epochs_proto.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

PRIMES = [17, 29, 37, 53, 67, 79, 97, 131, 193, 257, 389, 521, 769, 1031, 1543, 2053, 3079, 6151, 12289, 24593, 49157, 98317,
          196613, 393241, 786433, 1572869, 3145739, 6291469, 12582917, 25165843, 50331653, 100663319, 201326611, 402653189,
          805306457, 1610612741, 3221225473, 4294967291]


def prime_at_least(n):
    for p in PRIMES:
        if p >= n:
            return p
    return PRIMES[-1]


def grow_to(count):
    return prime_at_least(max(count + 1, count + (count >> 1)) + 1)


def epoch_order(iv, rank, H, W, gap, long_enough):
    """iv: [n][4] (a, b, startWord, endWord); rank: [W][H] seed-group ranks; returns interval indices in emission order."""
    a, b, s, e = (iv[:, i].astype(np.int64) for i in range(4))
    n = len(iv)
    # creation order: (start word, rank of a's word group, a, b)
    created = np.lexsort((b, a, rank[s, a], s))
    a, b, s, e = a[created], b[created], s[created], e[created]
    key = a * H + b
    flush = np.minimum(e + gap + 1, W)  # phase after whose inserts the node leaves; W = the final flush
    start_begin = np.searchsorted(s, np.arange(W + 1))
    ends = np.bincount(e, minlength=W)
    # rehash schedule (count only)
    epochs = [(0, 17)]  # (first creation rank, bucket count)
    count, B = 0, 17
    for w in range(W):
        q, q_end = start_begin[w], start_begin[w + 1]
        while q < q_end:
            room = B - count  # inserts that fit before count + 1 > B
            if q + room >= q_end:
                count += q_end - q
                q = q_end
            else:
                q += room
                count += room
                want = grow_to(count)
                if want != B:
                    B = want
                    epochs.append((int(q), B))
                else:  # cannot happen (grow_to(count) > count), kept for symmetry with the reference
                    count += 1
                    q += 1
        if w - gap - 1 >= 0:
            count -= ends[w - gap - 1]
    # epoch in which each flush phase happens: the last rehash at a rank < start_begin[f + 1]
    q_of_epoch = np.array([q for q, _ in epochs])
    phase_epoch = np.searchsorted(q_of_epoch, start_begin[np.minimum(np.arange(W + 1) + 1, W)], side="left") - 1
    phase_epoch[W] = len(epochs) - 1
    g = np.zeros(n, np.int64)
    wk = np.zeros(n, np.int64)
    fin_g = np.zeros(n, np.int64)
    fin_w = np.zeros(n, np.int64)
    for k, (qk, Bk) in enumerate(epochs):
        q_next = epochs[k + 1][0] if k + 1 < len(epochs) else n
        if k == 0:
            carried = np.zeros(0, np.int64)
            base = 1
        else:
            wk_word = s[qk]
            carried = np.flatnonzero((np.arange(n) < qk) & (e >= wk_word - gap - 1))
            N = len(carried)
            walk = carried[np.lexsort((-wk[carried], -g[carried]))]  # list order: g descending, then w descending
            nb = key[walk] % Bk
            # first walk index per new bucket
            order_b = np.lexsort((np.arange(N), nb))
            first = np.full(N, 0, np.int64)
            sb = nb[order_b]
            is_first = np.r_[True, sb[1:] != sb[:-1]]
            first_idx = np.maximum.accumulate(np.where(is_first, np.arange(N), 0))
            first[order_b] = order_b[first_idx]  # walk index of the first node of my bucket (order_b sorted by (bucket, walk idx))
            g[walk] = N - first
            wk[walk] = np.arange(N)
            base = N + 1
        new = np.arange(qk, q_next)
        tick = base + (new - qk)
        wk[new] = tick
        nodes = np.concatenate([carried, new])
        is_new = np.concatenate([np.zeros(len(carried), bool), np.ones(len(new), bool)])
        bk = key[nodes] % Bk
        order = np.lexsort((nodes, is_new, bk))
        cur_b, cur_g, cur_until = -1, 0, -1
        for i in order:  # the device does this walk with one thread per bucket
            q = nodes[i]
            if bk[i] != cur_b:
                cur_b, cur_until = bk[i], -1
            if not is_new[i]:
                cur_g = g[q]
                cur_until = max(cur_until, flush[q])
            else:
                if cur_until >= s[q]:
                    g[q] = cur_g
                    cur_until = max(cur_until, flush[q])
                else:
                    g[q] = wk[q]
                    cur_g, cur_until = wk[q], flush[q]
        done = nodes[phase_epoch[flush[nodes]] == k]
        fin_g[done], fin_w[done] = g[done], wk[done]
    cand = np.flatnonzero(long_enough(s, e))
    out = cand[np.lexsort((-fin_w[cand], -fin_g[cand], flush[cand]))]
    return created[out]


def main():
    import conftest
    from fastsmc_b200 import asmc, synth
    from test_host_layer import _brute_force_intervals
    for n_haps, n_sites, gap, min_m in [(240, 3200, 1, 0.0), (400, 1920, 0, 0.05), (150, 6400, 2, 0.4), (300, 2560, 1, 0.02)]:
        root = f"/tmp/order_proto/d{n_haps}"
        os.makedirs("/tmp/order_proto", exist_ok=True)
        synth.dataset(root, n_haps, n_sites, 3000 * n_sites, 1, 77 + n_haps, founders=6)
        p = asmc.DecodingParams()
        p.verbose = False
        p.inFileRoot, p.decodingQuantFile, p.outFileRoot = root, conftest.DQ_69, root + ".out"
        p.decodingModeString, p.foldData, p.usingCSFS, p.FastSMC, p.hashing = "array", True, True, True, True
        p.gap, p.min_m = gap, min_m
        p.validateParamsFastSMC()
        d = asmc.Data(p)
        W = d.sites // 64
        iv = _brute_force_intervals(np.array(d.hapBits), W, gap)
        slow = np.array(asmc.pyASMC.replayReferenceOrder(iv, d, gap, min_m, fast=False))
        rank = np.array(asmc.pyASMC.seedGroupRanks(d)).astype(np.int64)
        gen = np.array(d.geneticPositions, np.float32)
        L = len(gen)

        def long_enough(s, e):
            end = np.minimum(64 * e + 63, L - 1)
            return 100.0 * (gen[end].astype(np.float64) - gen[64 * s].astype(np.float64)) >= float(np.float32(min_m))
        mine = epoch_order(iv, rank, n_haps, W, gap, long_enough)
        ok = len(mine) == len(slow) and np.array_equal(mine, slow)
        print(n_haps, n_sites, gap, min_m, "intervals", len(iv), "candidates", len(slow), "OK" if ok else "MISMATCH")
        if not ok:
            bad = np.flatnonzero(mine[:min(len(mine), len(slow))] != slow[:min(len(mine), len(slow))])
            print("  first mismatch at", bad[:5], len(mine), len(slow))
            sys.exit(1)


if __name__ == "__main__":
    main()

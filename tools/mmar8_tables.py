"""Greedy search for the row -> warp tables of k_mmar8 (bildk_mmar2.cuh, Mmar2Rows<GT, 8>): eight warps with one or two tile rows each,
warps w and w + 4 share a scheduler; minimise the largest per-scheduler sum of the row costs GTC + GT - ti, at most 16 accumulator tiles per warp."""
import itertools, random
def best(GT, MX, iters=200000, seed=0):
    GTC = GT + MX
    cost = [GTC + GT - ti for ti in range(GT)]
    rows = list(range(GT))
    rnd = random.Random(seed)
    bestv, besta = None, None
    # warps 0..7, each <=2 rows; SMSP k = warps k, k+4
    for it in range(iters):
        rnd.shuffle(rows)
        # greedy: assign rows (in shuffled-then-sorted order) to warp with least SMSP load that has room
        order = sorted(rows, key=lambda r: -cost[r] + rnd.random()*2)
        w = [[] for _ in range(8)]
        ok = True
        for r in order:
            cands = [i for i in range(8) if len(w[i]) < 2]
            if not cands: ok = False; break
            def smsp(i): return sum(cost[x] for x in w[i % 4]) + sum(cost[x] for x in w[i % 4 + 4])
            i = min(cands, key=lambda i: (smsp(i), len(w[i]), rnd.random()))
            w[i].append(r)
        if not ok: continue
        if any(len(x) == 0 for x in w): continue
        loads = [sum(cost[x] for x in w[k]) + sum(cost[x] for x in w[k + 4]) for k in range(4)]
        nacc = max(sum(GT - r for r in x) for x in w)
        if nacc > 16: continue
        v = (max(loads), nacc)
        if bestv is None or v < bestv:
            bestv, besta = v, [sorted(x) for x in w]
    return bestv, besta, sum(cost)
for GT in (10, 11, 12, 14):
    for MX in (0, 1):
        v, a, tot = best(GT, MX, 30000)
        print(GT, MX, v, a, "avg", tot / 4)

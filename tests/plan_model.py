"""Test infrastructure: a plain-Python restatement of the read kernel's device-side planner (rmnet_b200/csrc/sched.cuh), same
integer arithmetic, same tie-breaks, same record / piece emission.  Two uses:

  * CPU tests run it over thousands of random bank states and check the plan's invariants (every KV tile of every
    (object, query tile, Cv half) covered exactly once, slot / chain / piece-list bounds) -- the rules themselves;
  * GPU tests compare the plan the device built (read back through rmnet_memory_read_plan_host) with this one, piece by
    piece -- the device code against the rules.

Nothing under rmnet_b200/ imports this module."""

G_DEFAULT = 148
MAX_SPLITS, KV_TILE, MAX_TILES, QT = 16, 64, 64, 128
MAX_RECORDS, MAX_SEGS, MAX_STEPS, FILL_STRIDE, MIN_CHUNK = 160, 64, 72, 16, 4
FILL_MIN_LOAD, FILL_MAX_LOAD, N_MARGINS = 12, 96, 3
INF = 0xFFFFFFFF
PREC_SPLIT3, PREC_SINGLE, PREC_MIXED = 0, 1, 2


def plan_cost(precision):
    tile = 2700 if precision == PREC_SPLIT3 else (1740 if precision == PREC_MIXED else 1490)
    return tile, 14000, 6000


def _ceil_div(a, b):
    return -(-a // b)


def plan_deal(nt, nqt, G, cost):
    tile, first, extra = cost
    n = len(nt)
    max_nt = max([nt[o] if nqt[o] > 0 else 0 for o in range(n)] + [0])
    c_min = max(1, _ceil_div(max_nt, MAX_SPLITS))
    best, best_c = INF, c_min
    for c in range(1, 65):
        if not (c >= c_min and c <= MAX_TILES and c <= max(max_nt, 1)):
            continue
        items, longest = 0, 0
        for o in range(n):
            if nt[o] > 0 and nqt[o] > 0:
                ns = _ceil_div(nt[o], c)
                items += ns * 2 * nqt[o]
                longest = max(longest, _ceil_div(nt[o], ns))
        rounds = _ceil_div(items, G)
        cyc = (rounds * longest * tile + first + ((rounds - 1) & 0xFFFFFFFF) * extra) & 0xFFFFFFFF
        v = min(cyc >> 6, 0x3FFFFFF) * 64 + (64 - c)
        if v < best:
            best, best_c = v, c
    c = best_c
    if max_nt > MAX_TILES * MAX_SPLITS:
        c = c_min
    ns = [(max(_ceil_div(nt[o], c), min(MAX_SPLITS, _ceil_div(nt[o], MAX_TILES))) if (nt[o] > 0 and nqt[o] > 0) else 0) for o in range(n)]
    ibase, acc = [], 0
    for o in range(n):
        ibase.append(acc)
        acc += ns[o] * 2 * nqt[o]
    ibase.append(acc)
    cost0 = 0xFFFFFFFE if best == INF else (best >> 6) * 64
    return cost0, ns, ibase


def plan_fill(nt, nqt, G, cost, m):
    """-> (cost or INF, ns, records); a record = (o, slot, t0, len, cta_begin, n_ctas, unit0, gw)."""
    tile, first, extra = cost
    n = len(nt)
    total = sum(2 * nqt[o] * nt[o] for o in range(n))
    if total < G * FILL_MIN_LOAD or total > G * FILL_MAX_LOAD or G > 1023:
        return INF, None, None
    ideal = (total * tile + G * first) // G
    Lv = ideal + (tile * 3 * (m + 2)) // 4
    keys = {}
    for o in range(min(n, 64)):
        if nqt[o] > 0 and nt[o] > 0:
            keys[o] = (min(nqt[o], 255) << 24) | (min(nt[o], 0x3FFFF) << 6) | (63 - o)
    segs = [[0, 0, G, 0]]                      # load, begin, len, pieces   (index = creation order)
    ns = [0] * n
    recs = []
    steps = maxload = max_np = 0
    while keys:
        o = max(keys, key=lambda k: keys[k])
        key = keys.pop(o)
        g = 2 * (key >> 24)
        if g > G or (key >> 24) >= 255:
            return INF, None, None
        rem, slot, t0 = nt[o], 0, 0
        while rem > 0:
            steps += 1
            if steps > MAX_STEPS:
                return INF, None, None
            need, taken, base, first_len = g, [], 0, 0
            live = sorted([i for i in range(len(segs)) if segs[i][2] > 0], key=lambda i: (segs[i][0], i))
            for i in live:
                if need == 0:
                    break
                k = min(need, segs[i][2])
                if need == g:
                    first_len = segs[i][2]
                taken.append(i)
                need -= k
                base = segs[i][0]
                last_seg, last_k = i, k
            if need > 0:
                return INF, None, None
            avail = Lv - base - (first if base == 0 else extra)
            room = avail // tile if avail > 0 else 0
            left = MAX_SPLITS - slot
            nb = 1
            if left == 1:
                ln = rem
            else:
                ln0 = min(max(room, MIN_CHUNK), MAX_TILES)
                ln = min(ln0, rem)
                lo = rem - MAX_TILES * (left - 1)
                if ln < lo:
                    ln = lo
                tail = rem - ln
                if 0 < tail < MIN_CHUNK:
                    ln = rem if rem <= min(MAX_TILES, room + MIN_CHUNK) else rem - MIN_CHUNK
                if ln == ln0 and first_len >= 2 * g:
                    nb = min(first_len // g, left - 1, rem // ln)
                    while nb >= 2:
                        rem_j, left_j = rem - (nb - 1) * ln, left - (nb - 1)
                        tail_j = rem_j - ln
                        if left_j >= 2 and rem_j - MAX_TILES * (left_j - 1) <= ln and (tail_j == 0 or tail_j >= MIN_CHUNK):
                            break
                        nb -= 1
                    nb = max(nb, 1)
                    if nb > 1:
                        last_k = g * nb
            if ln > MAX_TILES:
                return INF, None, None
            unit = 0
            for i in sorted(taken):                                   # pass 2 walks the taken segments in index order
                ld, bg, sl, np_ = segs[i]
                np_ += 1
                k = last_k if i == last_seg else sl
                nl = ld + (first if ld == 0 else extra) + ln * tile
                maxload, max_np = max(maxload, nl), max(max_np, np_)
                if len(recs) >= MAX_RECORDS or (k < sl and len(segs) >= MAX_SEGS) or nl >= (1 << 25) or np_ > FILL_STRIDE:
                    return INF, None, None
                recs.append((o, slot, t0, ln, bg, k, unit, g if nb > 1 else k))
                if k == sl:
                    segs[i] = [nl, bg, k, np_]
                else:
                    segs[i] = [ld, bg + k, sl - k, np_ - 1]
                    segs.append([nl, bg, k, np_])
                unit += k
            rem -= ln * nb
            t0 += ln * nb
            slot += nb
        ns[o] = slot
    if max_np > FILL_STRIDE:
        return INF, None, None
    return maxload, ns, recs


def build_plan(counts, q_cells, G=G_DEFAULT, precision=PREC_SPLIT3):
    """counts[o] stored cells, q_cells[o] cells of the query rectangle -> (winner, ns, per-CTA piece lists); a piece =
    (object, query_tile, half, slot, first_tile, tiles, stored_cells, live_rows), like MemoryBank.read_plan plus the rows."""
    n = len(counts)
    nt = [_ceil_div(c, KV_TILE) for c in counts]
    nqt = [_ceil_div(q, QT) for q in q_cells]
    cost = plan_cost(precision)
    cost0, ns0, ibase = plan_deal(nt, nqt, G, cost)
    win, best, ns, recs = 0, cost0, ns0, None
    for m in range(N_MARGINS):
        c, ns_m, recs_m = plan_fill(nt, nqt, G, cost, m)
        if c < best:
            win, best, ns, recs = 1 + m, c, ns_m, recs_m
    lists = [[] for _ in range(G)]

    def rows(o, qt):
        return min(QT, q_cells[o] - qt * QT)

    if win == 0:
        n_items = ibase[n]
        for c in range(G):
            for item in range(c, n_items, G):
                o = 0
                while ibase[o + 1] <= item:
                    o += 1
                r = item - ibase[o]
                q, qt = divmod(r, nqt[o])
                half, j = q & 1, q >> 1
                t0, t1 = (j * nt[o]) // ns[o], ((j + 1) * nt[o]) // ns[o]
                lists[c].append((o, qt, half, j, t0, t1 - t0, counts[o], rows(o, qt)))
    else:
        for (o, slot, t0, ln, bg, k, unit0, gw) in recs:
            for d in range(k):
                grp = d // gw
                unit = unit0 + (d - grp * gw)
                half, qt = divmod(unit, nqt[o])
                lists[bg + d].append((o, qt, half, slot + grp, t0 + grp * ln, ln, counts[o], rows(o, qt)))
    return win, ns, lists


def check_plan(ns, lists, counts, q_cells, strict_chain=True, max_pieces=None):
    """Invariants of a plan (device-built or modelled); pieces are tuples whose first seven fields are
    (object, query_tile, half, slot, first_tile, tiles, stored_cells)."""
    n = len(counts)
    cover = {}
    for c, pcs in enumerate(lists):
        if max_pieces is not None:
            assert len(pcs) <= max_pieces
        for p in pcs:
            o, qt, half, slot, t0, ln, cnt = p[:7]
            assert 0 <= o < n and half in (0, 1) and ln > 0 and t0 >= 0
            assert cnt == counts[o], "piece carries the object's stored-cell count"
            assert slot < ns[o] <= 16
            cover.setdefault((o, qt, half), []).append((t0, ln, slot, c))
    for o in range(n):
        nt = (counts[o] + 63) // 64
        nqt = (q_cells[o] + 127) // 128
        if nt == 0 or nqt == 0:
            assert ns[o] == 0 and not any(k[0] == o for k in cover)
            continue
        chunks0 = None
        for qt in range(nqt):
            for half in (0, 1):
                segs = sorted(cover.get((o, qt, half), []))
                assert segs, f"object {o} tile {qt} half {half} has no pieces"
                pos = 0
                for (t0, ln, slot, c) in segs:
                    assert t0 == pos, "KV tiles covered once, in order, without gaps"
                    pos += ln
                assert pos == nt
                assert sorted(s[2] for s in segs) == list(range(ns[o])), "every partial slot written exactly once"
                chunks = [(t0, ln, slot) for (t0, ln, slot, c) in segs]
                if chunks0 is None:
                    chunks0 = chunks
                assert chunks == chunks0, "all query tiles / halves of an object share the chunking (slot <-> chunk)"
                if strict_chain and nt <= 16 * 64:
                    assert max(ln for _, ln, _ in chunks) <= 64, "accumulation-chain bound"
        assert not any(k[0] == o and k[1] >= nqt for k in cover)

"""CPU prototype (design validation for round 2): the decoder's per-slice chain pass WITHOUT the serial loop.

Today `k_dec_chain` (ckl_decode.cu) walks the b / t events of a slice one after the other (0.76 ms at 1024^3: the
revisit stack makes event k+1 depend on event k).  This file states the same computation as scans, one stable sort by
stack level and binary searches, and checks it against the serial loop on random event sequences.

Serial semantics (crackcodes.hpp:706-862 as restated in ckl_decode.cu):
    per chain (start S from the BOC index, ends at the 't' that finds the stack empty):
        segstart[k] = pos;  pos += s[k]
        'b': push(pos)            't': if stack: pos = pop()
Closed form.  Let D[k] be the stack size before event k, m(t) the 'b' a 't' pops, parent(b) the stack top before a push.
A 'b' stops contributing to `pos` when its PARENT is popped, and the segment before a 't' never contributes, so
    pos_after[k] = S + sum_{b_j <= k} s_j  -  sum_{t_k' <= k, stack non-empty} R(k'),
    R(k') = sum of s over the direct children of m(k') = A_{L+1}(k') - A_{L+1}(m(k')),  L = D[m(k')],
with A_L the running sum of s over the level-L pushes.  Grouped by level (level of a 'b' = D, of a 't' = D - 1), events
alternate b, t, b, t ..., so m(t) is simply the previous element of t's level group.
Chain boundaries: with +1 / -1 per b / t the running sum reaches a NEW minimum exactly at the chain-ending 't'.

The reference's push quirk (x == sx is stored as (0, y + 1), crackcodes.hpp:772,850) is not linear; a slice in which any
pushed position has x == sx has to take the serial path (it cannot happen for encoder-made streams: a branch point never
lies on the image border)."""
import numpy as np


def serial(types, s, starts):
    """types: 1 = 'b', 0 = 't'; s: (n, 2) segment displacements; starts: chain start positions.  -> segstart (n, 2), used"""
    n = len(types)
    out = np.zeros((n, 2), np.int64)
    open_, ci, stack = 0, 0, []
    pos = np.zeros(2, np.int64)
    for k in range(n):
        if open_ == 0:
            if ci >= len(starts):
                return out[:k], k
            pos = np.array(starts[ci], np.int64); ci += 1
            open_, stack = 1, []
        out[k] = pos
        pos = pos + s[k]
        if types[k]:
            open_ += 1
            stack.append(pos.copy())
        else:
            open_ -= 1
            if stack:
                pos = stack.pop()
    return out, n


def parallel(types, s, starts):
    types = np.asarray(types, np.int64)
    s = np.asarray(s, np.int64)
    n = len(types)
    if n == 0:
        return np.zeros((0, 2), np.int64), 0
    step = np.where(types == 1, 1, -1)
    P = np.cumsum(step)                                         # inclusive running sum of +1 / -1
    runmin = np.minimum.accumulate(np.concatenate([[0], P]))    # running minimum before each event (incl. the initial 0)
    chain_end = P < runmin[:-1]                                 # new minimum: the 't' that found the stack empty
    chain_id = np.concatenate([[0], np.cumsum(chain_end)[:-1]]) # chains are numbered in order
    used = n
    if chain_id.max() >= len(starts):                           # more chains in the events than start points: stop there
        used = int(np.argmax(chain_id >= len(starts)))
        types, s, step, P, chain_end, chain_id = types[:used], s[:used], step[:used], P[:used], chain_end[:used], chain_id[:used]
        n = used
        if n == 0:
            return np.zeros((0, 2), np.int64), 0
    # stack size before each event: running sum relative to the chain's base (the base drops by one per finished chain)
    D = np.concatenate([[0], P[:-1]]) + chain_id
    level = np.where(types == 1, D, D - 1)                      # chain-ending 't' has level -1: no match
    # group by (chain, level), stable in event order
    order = np.lexsort((np.arange(n), level, chain_id))
    same_group_prev = np.zeros(n, bool)
    same_group_prev[1:] = (chain_id[order][1:] == chain_id[order][:-1]) & (level[order][1:] == level[order][:-1])
    prev_in_group = np.full(n, -1, np.int64)
    prev_in_group[order[1:]] = np.where(same_group_prev[1:], order[:-1], -1)
    is_t_pop = (types == 0) & ~chain_end
    match = np.where(is_t_pop, prev_in_group, -1)               # the 'b' a popping 't' returns to
    assert np.all(types[match[is_t_pop]] == 1)
    # A_L: running sum of s over the pushes of one (chain, level) group, looked up by binary search
    bidx = np.nonzero(types == 1)[0]
    key = chain_id[bidx] * (n + 2) + level[bidx]
    o2 = np.lexsort((bidx, key))
    gk, gi = key[o2], bidx[o2]                                  # sorted by (group key, index)
    gs = np.cumsum(s[gi], axis=0)
    gs0 = np.concatenate([np.zeros((1, 2), np.int64), gs])      # exclusive prefix over the sorted pushes
    def A(group_key, upto):
        """sum of s over pushes of `group_key` with index <= upto (arrays)"""
        comp = gk * (n + 1) + gi
        lo = np.searchsorted(comp, group_key * (n + 1), side="left")
        hi = np.searchsorted(comp, group_key * (n + 1) + upto, side="right")
        return gs0[hi] - gs0[lo]
    R = np.zeros((n, 2), np.int64)
    tk = np.nonzero(is_t_pop)[0]
    if len(tk):
        m = match[tk]
        child_key = chain_id[m] * (n + 2) + level[m] + 1
        R[tk] = A(child_key, tk) - A(child_key, m)
    contrib = np.where((types == 1)[:, None], s, 0) - R
    # per-chain inclusive prefix of the contributions
    tot = np.cumsum(contrib, axis=0)
    first = np.concatenate([[True], chain_id[1:] != chain_id[:-1]])
    base = np.zeros((n, 2), np.int64)
    fi = np.nonzero(first)[0]
    base_per_chain = np.concatenate([np.zeros((1, 2), np.int64), tot])[fi]      # total before the chain's first event
    pos_after = np.asarray(starts, np.int64)[chain_id] + tot - base_per_chain[chain_id]
    segstart = np.where(first[:, None], np.asarray(starts, np.int64)[chain_id], np.concatenate([np.zeros((1, 2), np.int64), pos_after[:-1]]))
    return segstart, used


def random_events(rng, n_chains, max_events):
    types = []
    for _ in range(n_chains):
        open_, depth, budget = 1, 0, int(rng.integers(1, max_events))
        while open_ > 0:
            if budget > 0 and (depth == 0 or rng.random() < 0.55):
                types.append(1); open_ += 1; depth += 1; budget -= 1
            else:
                types.append(0); open_ -= 1; depth = max(depth - 1, 0)
    return np.array(types, np.int64)


def main():
    rng = np.random.default_rng(0)
    cases = 0
    for trial in range(300):
        types = random_events(rng, int(rng.integers(1, 6)), int(rng.integers(2, 400)))
        n = len(types)
        s = rng.integers(-40, 41, size=(n, 2))
        n_chains = int(np.sum(np.cumsum(np.where(types == 1, 1, -1)) < np.minimum.accumulate(np.concatenate([[0], np.cumsum(np.where(types == 1, 1, -1))]))[:-1]))
        nstarts = n_chains if trial % 5 else max(n_chains - 1, 0)            # sometimes fewer start points than chains
        starts = [tuple(rng.integers(0, 1000, size=2)) for _ in range(nstarts)]
        a, ua = serial(types, s, starts)
        b, ub = parallel(types, s, starts)
        assert ua == ub, (trial, ua, ub)
        assert np.array_equal(a, b[:ua]), trial
        cases += 1
    print(f"chain scan prototype: {cases} random event sequences match the serial pass")


if __name__ == "__main__":
    main()

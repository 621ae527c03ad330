"""CPU prototype (design validation for round 2): parsing the order-N crack-code bitstream in parallel.

`k_dec_markov` (ckl_decode.cu) decodes a slice's bitstream symbol by symbol.  Two things are serial in it, and only one has
to be: (1) finding where each code starts, and (2) the context chain symbol -> context -> model row -> next symbol.
The code is 0 / 10 / 110 / 111 for ranks 0..3 (markov.hpp:444-458) and its LENGTHS do not depend on the model, so (1) is a
three-state automaton over the bits (state = number of leading ones seen: 0, 1, 2) and can be done with a scan:

    per 32-bit word:  T[s] = (state after the word, ranks emitted) for each entry state s
    exclusive scan of the state maps -> entry state of every word -> every word emits its ranks independently
    exclusive scan of the emitted counts -> where each word's ranks go

What stays serial is (2), one table lookup per symbol on ranks that are already unpacked.  The reference decodes while
pos < nbits and reads zeros past the end (markov.hpp:278-323), i.e. a code cut off by the end of the stream is completed
with zero bits; the first symbol is two raw bits.  Checked here against the symbol-by-symbol parse on random streams."""
import numpy as np


def serial_ranks(bits):
    """bits: array of 0/1 (LSB-first order of the stream), starting AFTER the two raw bits.  -> list of ranks"""
    n = len(bits)
    out = []
    pos = 0
    get = lambda i: int(bits[i]) if i < n else 0   # noqa: E731
    while pos < n:
        if get(pos) == 0:
            out.append(0); pos += 1
        elif get(pos + 1) == 0:
            out.append(1); pos += 2
        elif get(pos + 2) == 0:
            out.append(2); pos += 3
        else:
            out.append(3); pos += 3
    return out


def word_table(word_bits):
    """for each entry state: (exit state, ranks emitted inside the word)"""
    tab = []
    for s0 in range(3):
        s, ranks = s0, []
        for b in word_bits:
            if b == 0:
                ranks.append(s); s = 0             # 0 closes the code: rank = number of ones before it
            elif s == 2:
                ranks.append(3); s = 0             # third one: 111
            else:
                s += 1
        tab.append((s, ranks))
    return tab


def parallel_ranks(bits, W=32):
    n = len(bits)
    nw = (n + W - 1) // W
    tabs = [word_table(bits[i * W:(i + 1) * W]) for i in range(nw)]          # independent per word
    # exclusive scan of the state maps (composition is associative; 3 states -> a 6-bit map per word on the GPU)
    maps = [tuple(t[s][0] for s in range(3)) for t in tabs]
    entry = []
    cur = (0, 1, 2)
    for mp in maps:                                                          # (a scan; written as a loop for clarity)
        entry.append(cur[0])                                                 # the stream starts in state 0
        cur = tuple(mp[cur[s]] for s in range(3))
    counts = [len(tabs[i][entry[i]][1]) for i in range(nw)]
    offs = np.concatenate([[0], np.cumsum(counts)])
    out = np.zeros(int(offs[-1]) + 1, np.int64)
    for i in range(nw):                                                      # independent per word
        r = tabs[i][entry[i]][1]
        out[offs[i]:offs[i] + len(r)] = r
    total = int(offs[-1])
    final_state = cur[0]
    if final_state:                                                          # a cut-off code is completed with zero bits
        out[total] = final_state
        total += 1
    return list(out[:total])


def main():
    rng = np.random.default_rng(1)
    cases = 0
    for trial in range(400):
        n = int(rng.integers(0, 700))
        p1 = rng.uniform(0.1, 0.9)
        bits = (rng.random(n) < p1).astype(np.int64)
        assert serial_ranks(bits) == parallel_ranks(bits), trial
        cases += 1
    print(f"markov parse prototype: {cases} random bitstreams parse identically")


if __name__ == "__main__":
    main()

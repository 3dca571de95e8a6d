#!/usr/bin/env python
"""Bank-conflict model of the row kernels' line exchange buffer (exb_fft8.cuh, ExLine).

8-byte elements, 16 bank pairs; a warp-wide 8-byte access is served per half-warp, one wavefront if the 16
slots hit 16 different bank pairs.  `patterns(N)` lists every access of fft8_run / unpack_store for one line of
N points held by N/8 threads; `cost(slot, N)` returns (wavefronts, ideal).  Run as a script to compare the
XOR swizzle with the former `i + i/8` padding."""


def slot_swizzle(i):
    return i ^ ((i >> 4) & 7) ^ ((i >> 3) & 8)


def slot_padded(i):
    return i + (i >> 3)


def patterns(N):
    P = N // 8
    npass = 2 if N == 64 else (3 if N <= 512 else 4)
    pats = []
    for h in range(0, P, 16):
        lanes = list(range(h, min(h + 16, P)))
        for r in range(8):                                    # pass 1 stores: 8 j + r
            pats.append(("st1", [8 * j + r for j in lanes]))
        for q in range(8):                                    # loads after every exchange: j + P q
            pats.append(("ld", [j + P * q for j in lanes]))
        if npass >= 3:
            for r in range(8):                                # pass 2 stores: 64 m + k + 8 r
                pats.append(("st2", [(j - (j & 7)) * 8 + (j & 7) + 8 * r for j in lanes]))
        if npass == 4:
            for r in range(8):                                # pass 3 stores: 512 m + k + 64 r
                pats.append(("st3", [(j - (j & 63)) * 8 + (j & 63) + 64 * r for j in lanes]))
        for q in range(4):                                    # two-for-one split: partner N - k
            pats.append(("rev", [(N - (j + P * q)) % N for j in lanes]))
    return pats


def cost(slot, N, kinds=None):
    tot = ideal = 0
    for name, idx in patterns(N):
        if kinds is not None and name not in kinds:
            continue
        banks = {}
        for i in idx:
            b = slot(i) % 16
            banks[b] = banks.get(b, 0) + 1
        tot += max(banks.values())
        ideal += 1
    return tot, ideal


if __name__ == "__main__":
    for N in (128, 256, 512, 1024, 2048):
        print(N, "swizzle", cost(slot_swizzle, N), "padded", cost(slot_padded, N))

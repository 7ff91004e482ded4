"""CPU emulation of the index arithmetic of the round-2 kernel candidates (none of them has run on a GPU
yet): each check re-states the mapping the device code uses and compares it with the mapping it must be
consistent with.  python tools/emulate_candidates.py

  1. Gram micro-block walk (AB_GRAM_MICRO, gram_kernel.cuh decode_tile / advance_item): every lower tile
     is visited exactly once for any grid size.
  2. TMA GEMM (gemm_tma.cu): where SWIZZLE_128B puts element (row, k) of a 16 x 16 fp64 box == the
     kernel's fragment offsets (swz(), the per-k-step flip); bank-conflict degree of a fragment load.
  3. GEMM fast loader (AB_GEMM_FASTLOAD, gemm.cu FastTile) == load_tile's chunk mapping.
The accuracy of the scaled-domain exp (AB_GRAM_SCALEDEXP) is emulated in tools/exp_emulation.c (gcc -O2 -lm).
"""
import collections

import numpy as np

SB = 16


def decode(t, tiles_i, micro):
    mb = micro or 1
    if micro:
        mic, sub = t // (mb * mb), t % (mb * mb)
        sb, local = mic // (SB * SB), mic % (SB * SB)
    else:
        sb, local = t // (SB * SB), t % (SB * SB)
    i = int((np.sqrt(np.float32(8.0) * np.float32(sb) + np.float32(1.0)) - np.float32(1.0)) * np.float32(0.5))
    while i * (i + 1) // 2 > sb:
        i -= 1
    while (i + 1) * (i + 2) // 2 <= sb:
        i += 1
    if micro:
        I = mb * (i * SB + local % SB) + sub % mb
        J = mb * ((sb - i * (i + 1) // 2) * SB + local // SB) + sub // mb
    else:
        I, J = i * SB + local % SB, (sb - i * (i + 1) // 2) * SB + local // SB
    return (I < tiles_i and J <= I), I, J


def items(tiles_i, micro):
    if micro:
        nsb = ((tiles_i + micro - 1) // micro + SB - 1) // SB
        return nsb * (nsb + 1) // 2 * SB * SB * micro * micro
    nsb = (tiles_i + SB - 1) // SB
    return nsb * (nsb + 1) // 2 * SB * SB


def check_walk():
    for tiles_i in (1, 2, 3, 5, 16, 17, 31, 32, 33, 100, 512):
        for G in (1, 7, 296):
            for micro in (0, 2, 4):  # micro-block edge MB (0 = plain order)
                n = items(tiles_i, micro)
                per = micro * micro if micro else 1
                grid = min(n // per, G)
                seen = set()
                for b in range(grid):
                    t = per * b
                    while t < n:
                        ok, I, J = decode(t, tiles_i, micro)
                        if ok:
                            assert (I, J) not in seen
                            seen.add((I, J))
                        t = (t + 1 if t % per != per - 1 else t + per * grid - (per - 1)) if micro else t + grid
                assert seen == {(I, J) for I in range(tiles_i) for J in range(I + 1)}
    print("1. tile walks cover the lower triangle exactly once (plain and micro-block order)")


def check_swizzle():
    box = 2048

    def tma(r, kk):
        lin = kk * 128 + (r & 15) * 8
        chunk = ((lin >> 4) & 7) ^ ((lin >> 7) & 7)
        return (r >> 4) * box + ((lin & ~0x70) | (chunk << 4))

    def swz(r, kk):
        ii = r & 15
        return (r >> 4) * box + kk * 128 + ((((ii >> 1) ^ (kk & 7)) << 4) | ((ii & 1) << 3))

    worst = 0
    for r in range(128):
        for kk in range(16):
            assert tma(r, kk) == swz(r, kk)
        for lr in range(4):
            for ks in range(4):
                assert (swz(r, lr) ^ (64 if ks & 1 else 0)) + ks * 512 == swz(r, 4 * ks + lr)
    for base in range(0, 128, 8):
        for ks in range(4):
            for half in range(2):
                banks = collections.Counter()
                for lane in range(16 * half, 16 * half + 16):
                    banks[(swz(base + (lane >> 2), 4 * ks + (lane & 3)) >> 3) & 15] += 1
                worst = max(worst, max(banks.values()))
    print(f"2. TMA swizzle == fragment offsets; worst bank multiplicity of a fragment load: {worst}-way")


def check_fast_loader():
    BK, T = 16, 256
    LDK = BK + 4
    for EXT in (128, 64):
        for km in (False, True):
            for tid in range(T):
                for it in range((BK * EXT // 2) // T):
                    chunk = tid + it * T
                    if not km:
                        want = (chunk // (EXT // 2)) * (EXT + 4) + (chunk % (EXT // 2)) * 2
                        kstep = T // (EXT // 2)
                        got = (tid // (EXT // 2)) * (EXT + 4) + (tid % (EXT // 2)) * 2 + it * kstep * (EXT + 4)
                    else:
                        want = (chunk // (BK // 2)) * LDK + (chunk % (BK // 2)) * 2
                        got = (tid // (BK // 2)) * LDK + (tid % (BK // 2)) * 2 + it * (T // (BK // 2)) * LDK
                    assert want == got
    print("3. FastTile chunk mapping == load_tile chunk mapping")


if __name__ == "__main__":
    check_walk()
    check_swizzle()
    check_fast_loader()

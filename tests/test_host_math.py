"""CPU suite, part 3: host-side checks of the arithmetic identities the kernels rely on."""
import ctypes
import ctypes.util

import numpy as np

libm = ctypes.CDLL(ctypes.util.find_library("m"))
libm.fmaf.restype = ctypes.c_float
libm.fmaf.argtypes = [ctypes.c_float] * 3


def div3_fast(x):
    """common.cuh div3_exact(): q = x*RN(1/3); r = fma(-3,q,x); q' = fma(r,RN(1/3),q)"""
    c3 = np.float32(float.fromhex("0x1.555556p-2"))
    q = np.float32(x) * c3
    r = libm.fmaf(-3.0, q, x)
    return np.float32(libm.fmaf(r, c3, q))


def test_div3_fast_path_is_correctly_rounded_on_samples():
    # the exhaustive 2^32 sweep is done once in C; this samples every binade + random mantissas
    rng = np.random.default_rng(0)
    bits = np.concatenate([
        (np.arange(1, 255, dtype=np.uint32)[:, None] << 23 | rng.integers(0, 1 << 23, (254, 40), dtype=np.uint32)).ravel(),
        rng.integers(1, 1 << 23, 2000, dtype=np.uint32),                      # denormals
        np.array([0, 1, 0x007FFFFF, 0x00800000, 0x7F7FFFFF], dtype=np.uint32)])
    xs = bits.view(np.float32)
    for x in np.concatenate([xs, -xs[:2000]]):
        ref = np.float32(x) / np.float32(3.0)
        got = div3_fast(float(x))
        assert got == ref or (got == 0 and ref == 0), (float(x).hex(), float(ref).hex(), float(got).hex())


def test_minconv_backward_pass_identity():
    """aggregate.cu minconv_pair: with c >= 0 the reference's backward pass over the forward result F equals
    min(F, Bp), Bp = the same recurrence run downwards on the original values -- bit for bit in fp32."""
    rng = np.random.default_rng(1)
    for trial in range(200):
        n = int(rng.integers(2, 70))
        M = (rng.random(n) * rng.choice([1, 30, 1000])).astype(np.float32)
        if trial % 3 == 0:
            M[rng.random(n) < 0.3] = np.inf
        c = np.float32(rng.choice([0.0, 2.0, 8.0, 1.3, 0.1]))
        F = M.copy()
        for o in range(1, n):
            F[o] = min(np.float32(F[o - 1] + c), F[o])
        ref = F.copy()
        for o in range(n - 2, -1, -1):
            ref[o] = min(np.float32(ref[o + 1] + c), ref[o])
        Bp = M.copy()
        for o in range(n - 2, -1, -1):
            Bp[o] = min(np.float32(Bp[o + 1] + c), Bp[o])
        assert np.array_equal(np.minimum(F, Bp), ref)


def test_monotone_rounded_add_distributes_over_min():
    """RN(min(a,b)+c) == min(RN(a+c), RN(b+c)): why the neighbour-side transform can be computed once by the
    producer pixel and why padded +INF labels never leak into real labels."""
    rng = np.random.default_rng(2)
    a = (rng.random(100000) * 100).astype(np.float32)
    b = (rng.random(100000) * 100).astype(np.float32)
    c = np.float32(2.0)
    assert np.array_equal(np.minimum(a, b) + c, np.minimum(a + c, b + c))


def test_minconv_second_half_runs_on_the_partner_partials_only():
    """aggregate.cu minconv_half: Q = min(F, B) satisfies Q[o] = min(Q[o+1]+c, F[o]) below the meeting point and
    Q[o] = min(Q[o-1]+c, B[o]) above it, so the second half of each lane never re-reads the source vector."""
    rng = np.random.default_rng(3)
    for trial in range(300):
        n = 2 * int(rng.integers(1, 40))
        M = (rng.random(n) * rng.choice([1, 30, 1000])).astype(np.float32)
        if trial % 3 == 0:
            M[rng.random(n) < 0.3] = np.inf
        if trial % 7 == 0:
            M = np.round(M)
        c = np.float32(rng.choice([0.0, 2.0, 8.0, 1.3, 0.1]))
        F = M.copy()
        for o in range(1, n):
            F[o] = min(np.float32(F[o - 1] + c), F[o])
        B = M.copy()
        for o in range(n - 2, -1, -1):
            B[o] = min(np.float32(B[o + 1] + c), B[o])
        ref = np.minimum(F, B)
        h = n // 2
        Q = np.empty(n, np.float32)
        run = F[h - 1]                      # upward lane: carry of its first half, then over the partner's B[h:]
        for o in range(h, n):
            run = min(np.float32(run + c), B[o])
            Q[o] = run
        run = B[h]                          # downward lane: carry of its first half, then over the partner's F[:h]
        for o in range(h - 1, -1, -1):
            run = min(np.float32(run + c), F[o])
            Q[o] = run
        assert np.array_equal(Q, ref), (trial, n, float(c))


def test_finish_tile_band_bounds():
    """Fused finish (aggregate.cu tile_ready): the bands of a sweep that hold pixels of a rectangular tile are bounded
    by the band indices of two opposite corners in SCAN coordinates, because the scan coordinates are affine in
    (x, y) with slopes +-1 and the band index is monotone in them.  Restated here with the sweep table of
    common.cuh (pass_geometry) and checked against the brute-force set of bands over every pixel of the tile."""
    rng = np.random.default_rng(3)
    for _ in range(200):
        nx, ny = int(rng.integers(1, 70)), int(rng.integers(1, 70))
        T0, T1 = int(rng.integers(1, 20)), int(rng.integers(1, 20))
        shear = bool(rng.integers(0, 2))
        tw, th = int(rng.integers(1, 25)), int(rng.integers(1, 25))
        tx, ty = int(rng.integers(0, (nx + tw - 1) // tw)), int(rng.integers(0, (ny + th - 1) // th))
        x0, y0 = tx * tw, ty * th
        x1, y1 = min(x0 + tw, nx) - 1, min(y0 + th, ny) - 1
        for p in range(8):
            rm, incx, incy = (0x53 >> p) & 1, (0xC5 >> p) & 1, (0x99 >> p) & 1
            maxii, maxjj = (nx, ny) if rm else (ny, nx)

            def band(x, y):
                ax = x if incx else nx - 1 - x
                ay = y if incy else ny - 1 - y
                xs, ys = (ax, ay) if rm else (ay, ax)
                assert 0 <= xs < maxii and 0 <= ys < maxjj
                return (xs + ys) // T1 if (p >= 4 and shear) else ys // (T0 if p < 4 else T1)
            # the corner formula of tile_ready
            ax0, ax1 = (x0, x1) if incx else (nx - 1 - x1, nx - 1 - x0)
            ay0, ay1 = (y0, y1) if incy else (ny - 1 - y1, ny - 1 - y0)
            xs0, xs1, ys0, ys1 = (ax0, ax1, ay0, ay1) if rm else (ay0, ay1, ax0, ax1)
            if p >= 4 and shear:
                b0, b1 = (xs0 + ys0) // T1, (xs1 + ys1) // T1
            else:
                T = T0 if p < 4 else T1
                b0, b1 = ys0 // T, ys1 // T
            brute = {band(x, y) for x in range(x0, x1 + 1) for y in range(y0, y1 + 1)}
            assert min(brute) == b0 and max(brute) == b1, (p, nx, ny, shear)
            nb = (maxii + maxjj - 1 + T1 - 1) // T1 if (p >= 4 and shear) else (maxjj + (T0 if p < 4 else T1) - 1) // (T0 if p < 4 else T1)
            assert 0 <= b0 <= b1 < nb

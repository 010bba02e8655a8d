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


# ------------------------------------------------------------------------------------------ image files of the CLI
def test_image_reader_rejects_malformed_files(tmp_path):
    """mgm_b200/host/imgio.hpp validates every offset and length it takes from a file: truncated and malformed inputs
    end in an error message (exit code 3 of the CLI), never in an out-of-bounds read.  Built with -fsanitize=address
    so that a missed check fails the test."""
    import os, struct, subprocess, zlib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "imgio_probe")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address", "-I", os.path.join(root, "mgm_b200", "host"),
                           "-I", os.path.join(root, "include"), "-o", exe, os.path.join(root, "tests", "imgio_probe.cc"),
                           "-L", os.path.join(root, "mgm_b200"), "-lmgmb200", "-Wl,-rpath," + os.path.join(root, "mgm_b200"), "-lz"])

    def probe(data, name="f", out=None):
        path = str(tmp_path / name)
        open(path, "wb").write(data)
        r = subprocess.run([exe, path] + ([out] if out else []), capture_output=True, text=True, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0"))
        assert "AddressSanitizer" not in r.stderr, r.stderr[-2000:]
        assert r.returncode in (0, 3), (r.returncode, r.stderr[-500:])
        return r.stdout.strip()

    img = (np.arange(6 * 5 * 3) % 251).astype(np.uint8).reshape(5, 6, 3)
    # good files of every format
    pnm = b"P6\n6 5\n255\n" + img.tobytes()
    assert probe(pnm) == "ok 6 5 3"
    pfm = b"Pf\n6 5\n-1\n" + np.arange(30, dtype="<f4").tobytes()
    assert probe(pfm) == "ok 6 5 1"
    import io as _io
    bio = _io.BytesIO(); np.save(bio, img.astype(np.float32)); npy = bio.getvalue()
    assert probe(npy) == "ok 6 5 3"
    def png_bytes(w, h, raw, ctype=2, depth=8):
        def chunk(t, d): return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))
        return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b"")
    raw = b"".join(b"\x00" + img[y].tobytes() for y in range(5))
    png = png_bytes(6, 5, raw)
    assert probe(png) == "ok 6 5 3"
    def tiff_bytes(w, h, data, strip_off=None, strip_cnt=None, nent_extra=0):
        ents = [(256, 4, 1, w), (257, 4, 1, h), (258, 3, 1, 32), (259, 3, 1, 1), (273, 4, 1, 8 if strip_off is None else strip_off),
                (277, 3, 1, 1), (279, 4, 1, len(data) if strip_cnt is None else strip_cnt), (339, 3, 1, 3)]
        ifd = 8 + len(data)
        b = b"II*\x00" + struct.pack("<I", ifd) + data + struct.pack("<H", len(ents) + nent_extra)
        for t, ty, c, v in ents:
            b += struct.pack("<HHII", t, ty, c, v)
        return b + struct.pack("<I", 0)
    tif = tiff_bytes(6, 5, np.arange(30, dtype="<f4").tobytes())
    assert probe(tif) == "ok 6 5 1"
    # truncations of every format at every few bytes
    for good, okline in ((pnm, "ok 6 5 3"), (pfm, "ok 6 5 1"), (npy, "ok 6 5 3"), (png, "ok 6 5 3"), (tif, "ok 6 5 1")):
        for cut in list(range(0, min(len(good), 160), 3)) + [len(good) - 1, len(good) - 7]:
            r = probe(good[:cut])
            # an error, or -- when only trailing bytes that carry no pixels are missing (PNG IEND) -- the right image
            assert r.startswith("error") or (r == okline and cut > len(good) - 16), (cut, r)
    # malformed headers
    assert probe(b"P6\n60000 50000\n255\n" + img.tobytes()).startswith("error")
    assert probe(b"Pf\n-6 5\n-1\n" + pfm[10:]).startswith("error")
    assert probe(npy.replace(b"(5, 6, 3)", b"(5, 6, 9)")).startswith("error")
    assert probe(png_bytes(6, 50, raw)).startswith("error")                       # fewer scanlines than the header says
    assert probe(png_bytes(6, 5, raw, ctype=5)).startswith("error")
    assert probe(png[:33] + struct.pack(">I", 0x7fffffff) + png[37:]).startswith("error")   # chunk length beyond the file
    assert probe(tiff_bytes(6, 5, np.arange(30, dtype="<f4").tobytes(), strip_off=1 << 30)).startswith("error")
    assert probe(tiff_bytes(6, 5, np.arange(30, dtype="<f4").tobytes(), strip_cnt=1 << 30)).startswith("error")
    assert probe(tiff_bytes(6, 5, np.arange(30, dtype="<f4").tobytes(), nent_extra=4000)).startswith("error")
    assert probe(tif[:4] + struct.pack("<I", 1 << 31) + tif[8:]).startswith("error")   # directory offset beyond the file
    # writers: extensions this build cannot encode are refused, PFM only with 1 or 3 channels
    assert probe(pnm, out=str(tmp_path / "o.png")).startswith("error")
    assert probe(pnm, out=str(tmp_path / "o.npy")) == "ok 6 5 3" and np.load(str(tmp_path / "o.npy")).shape == (5, 6, 3)
    assert probe(pnm, out=str(tmp_path / "o.pfm")) == "ok 6 5 3"
    two = _io.BytesIO(); np.save(two, img[:, :, :2].astype(np.float32))
    assert probe(two.getvalue(), out=str(tmp_path / "o2.pfm")).startswith("error")
    assert probe(two.getvalue(), out=str(tmp_path / "o2.tif")) == "ok 6 5 2"

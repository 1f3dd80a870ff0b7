"""C++ host shim (scenelib2_b200/host): the MonoSLAM / Kalman / Feature class surface over the
C ABI, driven by the headless analogue of examples/MonoSlamSceneLib1.cpp."""
import os
import subprocess

import numpy as np
import pytest

from scenelib2_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "scenelib2_b200", "host")
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _write_case(tmp, sc, Pxx):
    """cfg in the reference's `key = value;` format (data/SceneLib2.cfg) + PGM templates + raw frames."""
    synth.write_reference_case(tmp, sc, Pxx)


def test_shim_builds_and_refuses_without_gpu(tmp_path):
    import __graft_entry__ as g
    g.build()
    exe = os.path.join(HOST, "sl2_headless")
    assert os.path.exists(exe) and os.path.exists(os.path.join(HOST, "libscenelib2_b200_host.so"))
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    kp = np.load(os.path.join(G, "known_patches.npy"))
    sc = synth.make_scene("C1", n_frames=1, known_patches=kp)
    _write_case(str(tmp_path), sc, np.eye(13) * 1e-4)
    r = subprocess.run([exe, str(tmp_path / "case.cfg"), str(tmp_path / "frames.raw"), "320", "240", "1"],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_headless_matches_oracle(tmp_path, oracle):
    from gpu_util import assert_state_close, oracle_slam_from_scene
    kp = np.load(os.path.join(G, "known_patches.npy"))
    sc = synth.make_scene("C1", n_frames=8, known_patches=kp)
    Pxx = np.diag([4e-4] * 3 + [2e-5] * 4 + [1e-3] * 3 + [1e-3] * 3)     # cf. data/SceneLib2.cfg:85-115
    sc.P0 = np.zeros_like(sc.P0)
    sc.P0[:13, :13] = Pxx
    _write_case(str(tmp_path), sc, Pxx)
    exe = os.path.join(HOST, "sl2_headless")
    out = tmp_path / "out.txt"
    r = subprocess.run([exe, str(tmp_path / "case.cfg"), str(tmp_path / "frames.raw"), "320", "240", "8",
                        str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    vals = np.loadtxt(str(out))
    n = int(vals[0])
    xg, Pg = vals[1:1 + n], vals[1 + n:].reshape(n, n).T
    o = oracle_slam_from_scene(oracle, sc)
    for t in range(8):
        o.step(sc.frames[t])
    xo, Po = o.get_state()
    assert n == o.n
    d = np.sqrt(np.abs(np.diag(Po))) + 1e-12
    assert (np.abs(Pg - Po) <= 1e-6 * d[:, None] * d[None, :] + 1e-18).all()
    assert np.allclose(xg, xo, rtol=1e-7, atol=1e-10)
    assert "measured 10" in r.stdout


_GRABBER_DRIVER = r"""
#include <chrono>
#include <cstdio>
#include <thread>
#include "scenelib2_b200.h"
int main(int argc, char **argv) {
  try {
    SceneLib2::FrameGrabber g;
    g.Init(argv[1], argc > 2);
    SceneLib2::Frame f;
    int id = 0;
    while (!g.Exhausted()) {
      if (!g.GetFrame(id, &f)) { std::this_thread::sleep_for(std::chrono::milliseconds(1)); continue; }
      unsigned long long sum = 0;
      for (int i = 0; i < f.data.rows * f.data.cols; ++i) sum = sum * 31 + f.data.data[i];
      std::printf("%d %d %d %llu\n", f.frame_id, f.data.cols, f.data.rows, sum);
      ++id;
    }
  } catch (const std::exception &e) { std::fprintf(stderr, "error: %s\n", e.what()); return 1; }
  return 0;
}
"""


def _checksum(a):
    s = 0
    for v in a.reshape(-1).tolist():
        s = (s * 31 + v) & 0xFFFFFFFFFFFFFFFF
    return s


def test_frame_grabber_reads_sorted_pgm_files(tmp_path):
    """N4: FrameGrabber / FileGrabber (framegrabber.cpp:59-103, filegrabber.cpp:54-110): recursive sorted file
    list, reader thread + bounded queue, frames in order; P5 and P2 PGMs with header comments decode, a
    non-image yields an empty frame, a missing directory throws, the USB mode is refused."""
    import __graft_entry__ as g
    g.build()
    rng = np.random.default_rng(4)
    d = tmp_path / "seq"
    (d / "sub").mkdir(parents=True)
    frames = [rng.integers(0, 256, (12, 20), dtype=np.uint8) for _ in range(60)]   # > queue bound of 50
    expect = {}
    for i, fr in enumerate(frames):
        name = (d / "sub" / ("b%04d.pgm" % i)) if i % 7 == 3 else (d / ("a%04d.pgm" % i))
        if i % 5 == 0:
            body = b"P2\n# ascii frame\n20 12\n255\n" + b" ".join(b"%d" % v for v in fr.reshape(-1)) + b"\n"
        else:
            body = b"P5\n# comment line\n20 12\n255\n" + fr.tobytes()
        name.write_bytes(body)
        expect[str(name)] = fr
    (d / "notes.txt").write_text("not an image")
    expect[str(d / "notes.txt")] = np.zeros((0, 0), np.uint8)
    src = tmp_path / "drv.cpp"
    src.write_text(_GRABBER_DRIVER)
    exe = tmp_path / "drv"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", "-I" + HOST, "-o", str(exe), str(src),
                           "-L" + HOST, "-lscenelib2_b200_host", "-L" + os.path.dirname(HOST), "-lsl2b200",
                           "-Wl,-rpath," + HOST + ":" + os.path.dirname(HOST)])
    r = subprocess.run([str(exe), str(d)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    rows = [ln.split() for ln in r.stdout.strip().splitlines()]
    order = sorted(expect)                                     # std::sort of the full path names
    assert [int(x[0]) for x in rows] == list(range(len(order)))
    for row, path in zip(rows, order):
        fr = expect[path]
        assert (int(row[1]), int(row[2])) == (fr.shape[1] if fr.size else 0, fr.shape[0] if fr.size else 0)
        assert int(row[3]) == _checksum(fr)
    r = subprocess.run([str(exe), str(tmp_path / "nowhere")], capture_output=True, text=True)
    assert r.returncode == 1 and "doesn't exist" in r.stderr   # filegrabber.cpp:81
    r = subprocess.run([str(exe), str(d), "usb"], capture_output=True, text=True)
    assert r.returncode == 1 and "USB" in r.stderr


@pytest.mark.gpu
def test_headless_directory_mode_equals_raw_mode(tmp_path):
    """sl2_headless fed by the FrameGrabber from a directory of PGM frames == the raw-file mode."""
    kp = np.load(os.path.join(G, "known_patches.npy"))
    sc = synth.make_scene("C1", n_frames=6, known_patches=kp)
    Pxx = np.diag([4e-4] * 3 + [2e-5] * 4 + [1e-3] * 3 + [1e-3] * 3)
    _write_case(str(tmp_path), sc, Pxx)
    d = tmp_path / "frames"
    d.mkdir()
    for t in range(6):
        (d / ("rawoutput%04d.pgm" % t)).write_bytes(b"P5\n320 240\n255\n" + sc.frames[t].tobytes())
    exe = os.path.join(HOST, "sl2_headless")
    a, b = tmp_path / "a.txt", tmp_path / "b.txt"
    r = subprocess.run([exe, str(tmp_path / "case.cfg"), str(tmp_path / "frames.raw"), "320", "240", "6",
                        str(a)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r2 = subprocess.run([exe, str(tmp_path / "case.cfg"), str(d), str(b)], capture_output=True, text=True)
    assert r2.returncode == 0, r2.stderr
    assert r.stdout == r2.stdout
    assert (np.loadtxt(str(a)) == np.loadtxt(str(b))).all()


def _png_bytes(img, color_type=0, depth=8, filters=None, level=6, palette=None, idat_split=None):
    """Minimal PNG writer (zlib from the standard library): img is (H, W[, C]) uint8/uint16 samples."""
    import struct
    import zlib
    h, w = img.shape[:2]
    ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[color_type]
    a = img.reshape(h, w, ch).astype(np.uint16 if depth == 16 else np.uint8)
    rows = []
    for y in range(h):
        if depth == 16:
            line = a[y].astype(">u2").tobytes()
        elif depth == 8:
            line = a[y].tobytes()
        else:
            bits = "".join(format(int(v), "0%db" % depth) for v in a[y].ravel())
            bits += "0" * (-len(bits) % 8)
            line = bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8))
        rows.append(np.frombuffer(line, np.uint8).astype(np.int32))
    bpp = max(1, ch * depth // 8)
    out, prev = bytearray(), np.zeros(len(rows[0]), np.int32)
    for y, cur in enumerate(rows):
        ft = 0 if filters is None else filters[y % len(filters)]
        left = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]])
        ul = np.concatenate([np.zeros(bpp, np.int32), prev[:-bpp]])
        if ft == 0:
            pred = 0
        elif ft == 1:
            pred = left
        elif ft == 2:
            pred = prev
        elif ft == 3:
            pred = (left + prev) // 2
        else:
            p = left + prev - ul
            pa, pb, pc = abs(p - left), abs(p - prev), abs(p - ul)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, ul))
        out.append(ft)
        out += ((cur - pred) & 255).astype(np.uint8).tobytes()
        prev = cur

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)
    z = zlib.compress(bytes(out), level)
    parts = [z] if not idat_split else [z[:idat_split], z[idat_split:]]
    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, color_type, 0, 0, 0))
    if palette is not None:
        png += chunk(b"PLTE", np.asarray(palette, np.uint8).tobytes())
    for p in parts:
        png += chunk(b"IDAT", p)
    return png + chunk(b"IEND", b"")


def test_png_and_pgm_frames_decode_like_imread(tmp_path):
    """FileGrabber::GetImageFile (filegrabber.cpp:106-109: cv::imread(path, 0)) on the shim: PNM and PNG -> 8-bit gray.
    Every PNG filter type, stored / fixed / dynamic deflate blocks, split IDAT, gray 1-16 bit, RGB(A), gray+alpha and
    palette images; the expected gray values are computed here with the formula cv::imread(path, 0) applies to a PNG
    (libpng's rgb-to-gray) and, where OpenCV is installed, compared with cv2.imread(path, 0) itself."""
    import ctypes as C
    import __graft_entry__ as g
    g.build()
    lib = C.CDLL(os.path.join(HOST, "libscenelib2_b200_host.so"))
    rng = np.random.default_rng(5)

    def decode(path):
        buf = np.zeros(1 << 20, np.uint8)
        w, h = C.c_int(0), C.c_int(0)
        rc = lib.sl2_host_decode_image(str(path).encode(), buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(w),
                                       C.byref(h))
        return rc, buf[:w.value * h.value].reshape(h.value, w.value) if rc == 0 else None

    def luma(rgb):   # libpng's png_set_rgb_to_gray(0.299, 0.587), which is what cv::imread(path, 0) uses for a PNG
        r, gr, b = [rgb[..., k].astype(np.int64) for k in range(3)]
        return ((r * 9797 + gr * 19234 + b * 3737) >> 15).astype(np.uint8)

    H, W = 37, 53
    smooth = (np.add.outer(np.arange(H) * 3, np.arange(W) * 2) % 256).astype(np.uint8)   # compressible: LZ77 matches
    noise = rng.integers(0, 256, (H, W), dtype=np.uint8)
    cases = []
    for name, img in (("smooth", smooth), ("noise", noise)):
        for level in (0, 1, 9):                       # 0 = stored blocks, 1 = mostly fixed, 9 = dynamic Huffman
            cases.append((name + "_l%d" % level, _png_bytes(img, filters=[0, 1, 2, 3, 4], level=level), img))
    cases.append(("split_idat", _png_bytes(smooth, filters=[4], idat_split=40), smooth))
    rgb = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    cases.append(("rgb", _png_bytes(rgb, color_type=2, filters=[1, 4, 3]), luma(rgb)))
    rgba = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    cases.append(("rgba", _png_bytes(rgba, color_type=6, filters=[2, 3]), luma(rgba)))
    ga = rng.integers(0, 256, (H, W, 2), dtype=np.uint8)
    cases.append(("gray_alpha", _png_bytes(ga, color_type=4, filters=[4]), ga[..., 0]))
    g16 = rng.integers(0, 65536, (H, W), dtype=np.uint16)
    cases.append(("gray16", _png_bytes(g16, depth=16, filters=[1, 2]), (g16 >> 8).astype(np.uint8)))
    for d in (1, 2, 4):
        v = rng.integers(0, 1 << d, (H, W), dtype=np.uint8)
        cases.append(("gray%d" % d, _png_bytes(v, depth=d), (v.astype(np.int32) * 255 // ((1 << d) - 1)).astype(np.uint8)))
    pal = rng.integers(0, 256, (16, 3), dtype=np.uint8)
    idx = rng.integers(0, 16, (H, W), dtype=np.uint8)
    cases.append(("palette4", _png_bytes(idx, color_type=3, depth=4, palette=pal), luma(pal[idx])))
    rgb16 = rng.integers(0, 65536, (H, W, 3), dtype=np.uint16)   # converted at 16 bits with rounding, then stripped
    r16, g16_, b16 = [rgb16[..., k].astype(np.int64) for k in range(3)]
    cases.append(("rgb16", _png_bytes(rgb16, color_type=2, depth=16, filters=[1, 3]),
                  (((r16 * 9797 + g16_ * 19234 + b16 * 3737 + 16384) >> 15) >> 8).astype(np.uint8)))
    try:
        import cv2   # the reference's own reader, when this image has it: the expected values above must be ITS values
    except ImportError:
        cv2 = None
    for name, data, want in cases:
        p = tmp_path / (name + ".png")
        p.write_bytes(data)
        rc, got = decode(p)
        assert rc == 0, name
        assert got.shape == want.shape and (got == want).all(), name
        if cv2 is not None:
            ref = cv2.imread(str(p), 0)
            assert ref is not None and ref.shape == got.shape and (ref == got).all(), name + " vs cv2.imread"
    # the PNM family (PBM / PGM / PPM, ASCII and binary, 8 and 16 bit) with the conversions OpenCV's reader applies
    def asc(v):
        return b" ".join(b"%d" % x for x in np.asarray(v).ravel()) + b"\n"

    def to8(v, mv, ascii_):
        v = np.asarray(v, np.int64)
        return v >> 8 if mv > 255 else (v * 255 // mv if (ascii_ and mv < 255) else v)

    h2, w2 = 13, 19
    for mv in (1, 15, 99, 255, 256, 1023, 65535):
        a = rng.integers(0, mv + 1, (h2, w2)).astype(np.uint16)
        c3 = rng.integers(0, mv + 1, (h2, w2, 3)).astype(np.uint16)
        raw = (lambda v: v.astype(">u2").tobytes()) if mv > 255 else (lambda v: v.astype(np.uint8).tobytes())
        pnm = [("p5", b"P5\n%d %d\n%d\n" % (w2, h2, mv) + raw(a), to8(a, mv, False)),
               ("p2", b"P2\n# comment\n%d %d\n%d\n" % (w2, h2, mv) + asc(a), to8(a, mv, True)),
               ("p6", b"P6\n%d %d\n%d\n" % (w2, h2, mv) + raw(c3), None), ("p3", b"P3\n%d %d\n%d\n" % (w2, h2, mv) + asc(c3), None)]
        for kind, data, want in pnm:
            if want is None:
                v = to8(c3, mv, kind == "p3")
                want = (v[..., 0] * 4899 + v[..., 1] * 9617 + v[..., 2] * 1868 + 8192) >> 14
            p = tmp_path / ("%s_%d.pnm" % (kind, mv))
            p.write_bytes(data)
            rc, got = decode(p)
            assert rc == 0 and (got == want).all(), p.name
            if cv2 is not None:
                assert (cv2.imread(str(p), 0) == got).all(), p.name + " vs cv2.imread"
    bits = rng.integers(0, 2, (h2, w2), dtype=np.uint8)
    for name, data in (("p1.pbm", b"P1\n%d %d\n" % (w2, h2) + asc(bits)),
                       ("p1_packed.pbm", b"P1\n%d %d\n" % (w2, h2) + b"".join(b"%d" % x for x in bits.ravel()) + b"\n"),
                       ("p4.pbm", b"P4\n%d %d\n" % (w2, h2) + np.packbits(bits, axis=1).tobytes())):
        (tmp_path / name).write_bytes(data)
        rc, got = decode(tmp_path / name)
        assert rc == 0 and (got == np.where(bits == 1, 0, 255)).all(), name
        if cv2 is not None:
            assert (cv2.imread(str(tmp_path / name), 0) == got).all(), name + " vs cv2.imread"
    # garbage and truncated files give "no image" like a failed imread
    (tmp_path / "a.pgm").write_bytes(b"P5\n%d %d\n255\n" % (W, H) + noise.tobytes())
    rc, got = decode(tmp_path / "a.pgm")
    assert rc == 0 and (got == noise).all()
    (tmp_path / "bad.png").write_bytes(cases[0][1][:60])
    (tmp_path / "text.txt").write_bytes(b"not an image")
    assert decode(tmp_path / "bad.png")[0] == -1 and decode(tmp_path / "text.txt")[0] == -1


def test_jpeg_frames_decode_like_imread(tmp_path):
    """FileGrabber::GetImageFile (filegrabber.cpp:106-109: cv::imread(path, 0)) on a JPEG: OpenCV asks libjpeg for
    grayscale output = the luminance component through libjpeg's slow-integer IDCT.  The shim's own decoder
    (host/jpeg_decode.h) must give the same bytes as cv2.imread(path, 0) -- the reference's very call -- for gray and
    colour files, every chroma sampling, optimised Huffman tables, restart intervals, odd sizes, OpenCV's and PIL's
    writers; progressive files are rejected like an unreadable image; truncated / corrupted files never crash."""
    cv2 = pytest.importorskip("cv2")
    Image = pytest.importorskip("PIL.Image")
    import ctypes as C
    import __graft_entry__ as g
    g.build()
    lib = C.CDLL(os.path.join(HOST, "libscenelib2_b200_host.so"))

    def decode(path):
        buf = np.zeros(1 << 20, np.uint8)
        w, h = C.c_int(0), C.c_int(0)
        rc = lib.sl2_host_decode_image(str(path).encode(), buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(w),
                                       C.byref(h))
        return rc, buf[:w.value * h.value].reshape(h.value, w.value) if rc == 0 else None

    rng = np.random.default_rng(11)
    checked = 0
    for (H, W) in ((1, 1), (7, 9), (16, 16), (17, 33), (240, 320)):
        yy, xx = np.mgrid[0:H, 0:W]
        base = ((np.sin(xx / 5.0) + np.cos(yy / 4.0)) * 50 + 128 + rng.normal(0, 12, (H, W))).clip(0, 255).astype(np.uint8)
        rgb = np.stack([base, np.roll(base, 1, 1), 255 - base], -1)
        files = []
        for sf in ("411", "420", "422", "440", "444"):
            p = tmp_path / ("cv_%d_%d_%s.jpg" % (H, W, sf))
            cv2.imwrite(str(p), rgb, [cv2.IMWRITE_JPEG_QUALITY, 80, cv2.IMWRITE_JPEG_SAMPLING_FACTOR,
                                      getattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR_" + sf)])
            files.append(p)
        for q, extra in ((10, []), (100, []), (75, [cv2.IMWRITE_JPEG_OPTIMIZE, 1]), (60, [cv2.IMWRITE_JPEG_RST_INTERVAL, 3])):
            p = tmp_path / ("cvg_%d_%d_%d_%d.jpg" % (H, W, q, len(extra)))
            cv2.imwrite(str(p), base, [cv2.IMWRITE_JPEG_QUALITY, q] + extra)
            files.append(p)
        for sub in (0, 1, 2):
            p = tmp_path / ("pil_%d_%d_%d.jpg" % (H, W, sub))
            Image.fromarray(rgb).save(str(p), quality=70, subsampling=sub, optimize=bool(sub & 1))
            files.append(p)
        for p in files:
            want = cv2.imread(str(p), 0)
            rc, got = decode(p)
            assert rc == 0 and got.shape == want.shape and (got == want).all(), p.name
            checked += 1
        p = tmp_path / ("prog_%d_%d.jpg" % (H, W))
        Image.fromarray(rgb).save(str(p), quality=70, progressive=True)
        assert decode(p)[0] == -1
    assert checked == 60
    data = (tmp_path / "cv_240_320_420.jpg").read_bytes()
    for cut in (3, 20, 200, len(data) // 2):
        (tmp_path / "cut.jpg").write_bytes(data[:cut])
        assert decode(tmp_path / "cut.jpg")[0] in (0, -1, -2)
    for k in range(200):
        b = bytearray(data)
        for _ in range(1 + k % 6):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        (tmp_path / "mut.jpg").write_bytes(bytes(b))
        assert decode(tmp_path / "mut.jpg")[0] in (0, -1, -2)    # -2: a mutated size field larger than the buffer


def test_save_patch_png_is_readable(tmp_path):
    """MonoSLAM::SavePatch writes the marked feature's template with cv::imwrite("patch.png", ...) (monoslam.cpp:1569).
    The shim's own PNG writer must produce a file any PNG reader accepts: read back with the shim's decoder and, where
    installed, with OpenCV and PIL (which check the chunk CRCs and the zlib Adler-32)."""
    import ctypes as C
    import __graft_entry__ as g
    g.build()
    lib = C.CDLL(os.path.join(HOST, "libscenelib2_b200_host.so"))
    rng = np.random.default_rng(2)
    for (h, w) in ((11, 11), (15, 15), (1, 1), (300, 400)):    # 300 x 400: more than one stored deflate block
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        p = tmp_path / ("patch_%d_%d.png" % (h, w))
        assert lib.sl2_host_write_png(str(p).encode(), img.ctypes.data_as(C.c_void_p), w, h) == 0
        buf = np.zeros(1 << 20, np.uint8)
        cw, ch = C.c_int(0), C.c_int(0)
        assert lib.sl2_host_decode_image(str(p).encode(), buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(cw),
                                         C.byref(ch)) == 0
        assert (buf[:w * h].reshape(h, w) == img).all()
        try:
            import cv2
            back = cv2.imread(str(p), cv2.IMREAD_UNCHANGED)
            assert back is not None and back.dtype == np.uint8 and (back == img).all()
        except ImportError:
            pass
        try:
            from PIL import Image
            with Image.open(str(p)) as im:
                im.load()
                assert im.mode == "L" and (np.asarray(im) == img).all()
        except ImportError:
            pass


def test_bmp_frames_decode_like_imread(tmp_path):
    """cv::imread(path, 0) on Windows bitmaps (filegrabber.cpp:106-109): 1 / 4 / 8 bit with a palette, 24 and 32 bit,
    bottom-up and top-down, from OpenCV's and PIL's writers.  The expected bytes are cv2.imread(path, 0)'s; the 32-bit
    case pins the single-precision formula OpenCV uses there (checked over all 2^24 colours when the decoder was written)."""
    cv2 = pytest.importorskip("cv2")
    Image = pytest.importorskip("PIL.Image")
    import ctypes as C
    import struct
    import __graft_entry__ as g
    g.build()
    lib = C.CDLL(os.path.join(HOST, "libscenelib2_b200_host.so"))

    def decode(path):
        buf = np.zeros(1 << 20, np.uint8)
        w, h = C.c_int(0), C.c_int(0)
        rc = lib.sl2_host_decode_image(str(path).encode(), buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(w),
                                       C.byref(h))
        return rc, buf[:w.value * h.value].reshape(h.value, w.value) if rc == 0 else None

    rng = np.random.default_rng(3)
    n = 0
    for (h, w) in ((11, 11), (13, 17), (240, 320)):
        gray = rng.integers(0, 256, (h, w), dtype=np.uint8)
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        files = []
        for name, arr in (("cv_g", gray), ("cv_c", rgb), ("cv_a", rgba)):
            p = tmp_path / ("%s_%d.bmp" % (name, w))
            cv2.imwrite(str(p), arr)
            files.append(p)
        for name, im in (("pil_L", Image.fromarray(gray)), ("pil_RGB", Image.fromarray(rgb)),
                         ("pil_1", Image.fromarray(gray).convert("1")),
                         ("pil_P", Image.fromarray(rgb).quantize(200)), ("pil_P4", Image.fromarray(rgb).quantize(16))):
            p = tmp_path / ("%s_%d.bmp" % (name, w))
            im.save(str(p), **({"bits": 4} if name == "pil_P4" else {}))
            files.append(p)
        rowb = (w * 3 + 3) // 4 * 4
        body = b"".join(rgb[y, :, ::-1].tobytes() + b"\0" * (rowb - 3 * w) for y in range(h))
        p = tmp_path / ("topdown_%d.bmp" % w)
        p.write_bytes(b"BM" + struct.pack("<IHHI", 54 + len(body), 0, 0, 54) +
                      struct.pack("<IiiHHIIiiII", 40, w, -h, 1, 24, 0, len(body), 2835, 2835, 0, 0) + body)
        files.append(p)
        for p in files:
            want = cv2.imread(str(p), 0)
            rc, got = decode(p)
            assert rc == 0 and want is not None and got.shape == want.shape and (got == want).all(), p.name
            n += 1
    assert n == 27
    data = (tmp_path / "cv_c_320.bmp").read_bytes()
    (tmp_path / "cut.bmp").write_bytes(data[:1000])
    assert decode(tmp_path / "cut.bmp")[0] == -1

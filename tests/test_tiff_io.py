"""hypelcnn_b200/utilities/tiff_io.py against libtiff as seen through Pillow and OpenCV (both directions, every layout
they support), hand-assembled files for what neither can write (big-endian, tiles, BigTIFF, planar, predictor 3), and
round trips of hyperspectral cubes (hundreds of samples per pixel)."""
import os
import struct
import zlib

import numpy
import pytest

from hypelcnn_b200.utilities.tiff_io import TiffError, imread, imwrite

RNG = numpy.random.default_rng(0)


def _cube(shape, dtype):
    if numpy.dtype(dtype).kind == "f":
        return (RNG.random(shape) * 100 - 50).astype(dtype)
    info = numpy.iinfo(dtype)
    return RNG.integers(info.min, int(info.max) + 1, shape).astype(dtype)


@pytest.mark.parametrize("shape,dtype,planar", [
    ((13, 17), numpy.uint8, "contig"), ((349, 190, 144), numpy.uint16, "contig"), ((5, 13, 17), numpy.float32, "separate"),
    ((9, 7, 3), numpy.uint8, "contig"), ((9, 7, 50), numpy.float64, "contig"), ((6, 5), numpy.int32, "contig"),
    ((12, 11, 65), numpy.float32, "contig"), ((4, 3, 2), numpy.int16, "separate"), ((1, 1), numpy.uint16, "contig")])
def test_round_trip(tmp_path, shape, dtype, planar):
    a = _cube(shape, dtype)
    path = str(tmp_path / "a.tif")
    imwrite(path, a, planarconfig=planar)
    b = imread(path)
    assert b.dtype == a.dtype and b.shape == a.shape and numpy.array_equal(a, b)
    assert numpy.array_equal(imread(open(path, "rb").read()), a)          # bytes are accepted as well


def test_bool_maps_and_rejections(tmp_path):
    path = str(tmp_path / "m.tif")
    imwrite(path, numpy.eye(5, dtype=bool))                                # shadow maps are written from bool arrays
    assert numpy.array_equal(imread(path), numpy.eye(5, dtype=numpy.uint8))
    for bad in (numpy.zeros((2, 2), numpy.float16), numpy.zeros((2, 2), numpy.complex64), numpy.zeros(4, numpy.uint8),
                numpy.zeros((2, 2, 2, 2), numpy.uint8)):
        with pytest.raises(TiffError):
            imwrite(path, bad)
    with pytest.raises(TiffError):
        imwrite(path, numpy.zeros((2, 2)), planarconfig="tiled")
    (tmp_path / "junk.tif").write_bytes(b"not a tiff at all")
    with pytest.raises(TiffError):
        imread(str(tmp_path / "junk.tif"))


def test_pillow_both_directions(tmp_path):
    from PIL import Image
    path = str(tmp_path / "p.tif")
    for a in (_cube((33, 41), numpy.uint8), _cube((33, 41), numpy.uint16), _cube((33, 41), numpy.float32),
              _cube((33, 41, 3), numpy.uint8)):
        imwrite(path, a)
        assert numpy.array_equal(numpy.array(Image.open(path)), a)                          # libtiff reads ours
        for compression in (None, "raw", "tiff_lzw", "tiff_adobe_deflate", "packbits"):
            Image.fromarray(a).save(path, **({"compression": compression} if compression else {}))
            b = imread(path)
            assert b.dtype == a.dtype and numpy.array_equal(a, b), compression            # we read libtiff's
    smooth = (numpy.add.outer(numpy.arange(300), numpy.arange(500)) // 3).astype(numpy.uint16)   # long LZW runs, > 64 KiB
    Image.fromarray(smooth).save(path, compression="tiff_lzw")
    assert numpy.array_equal(imread(path), smooth)


def test_opencv_both_directions(tmp_path):
    cv2 = pytest.importorskip("cv2")
    path = str(tmp_path / "c.tif")
    for a in (_cube((40, 50, 4), numpy.uint16), _cube((40, 50, 3), numpy.uint16), _cube((40, 50, 3), numpy.float32),
              _cube((40, 50), numpy.uint8), _cube((40, 50, 3), numpy.uint8),
              (numpy.arange(300 * 200).reshape(300, 200) % 5000).astype(numpy.uint16)):
        as_bgr = a if a.ndim == 2 else a[..., [2, 1, 0] + ([3] if a.shape[2] == 4 else [])]   # OpenCV's channel order
        assert cv2.imwrite(path, a)                     # LZW + horizontal predictor (integers) is OpenCV's default
        b = imread(path)
        assert b.dtype == a.dtype and numpy.array_equal(b, as_bgr)
        imwrite(path, a)
        assert numpy.array_equal(cv2.imread(path, cv2.IMREAD_UNCHANGED), as_bgr)


def _assemble(a, order="<", big=False, tile=None, planar=False, compression=1, predictor=1, pages=1):
    """A TIFF file built by hand: [H,W,S] array -> bytes with the requested byte order / tiling / layout."""
    height, width, samples = a.shape
    dt = a.dtype.newbyteorder(order)
    planes = [a[..., s:s + 1] for s in range(samples)] if planar else [a]
    blocks = []
    for plane in planes:
        if tile:
            th, tw = tile
            for y in range(0, height, th):
                for x in range(0, width, tw):
                    block = numpy.zeros((th, tw, plane.shape[2]), a.dtype)
                    part = plane[y:y + th, x:x + tw]
                    block[:part.shape[0], :part.shape[1]] = part
                    blocks.append(block)
        else:
            for y in range(0, height, 4):
                blocks.append(plane[y:y + 4])
    encoded = []
    for block in blocks:
        block = numpy.ascontiguousarray(block)
        if predictor == 2:
            block = numpy.concatenate([block[:, :1], numpy.diff(block, axis=1)], axis=1).astype(a.dtype)
        raw = block.astype(dt).tobytes()
        if predictor == 3:
            rows, w, s = block.shape
            be = block.astype(a.dtype.newbyteorder(">")).view(numpy.uint8).reshape(rows, w * s, a.dtype.itemsize)
            shuffled = numpy.ascontiguousarray(be.transpose(0, 2, 1)).reshape(rows, a.dtype.itemsize * w, s)
            diff = numpy.concatenate([shuffled[:, :1], numpy.diff(shuffled.astype(numpy.int16), axis=1).astype(numpy.uint8)], axis=1)
            raw = diff.astype(numpy.uint8).tobytes()
        encoded.append(zlib.compress(raw) if compression == 8 else raw)
    kind = {"u": 1, "i": 2, "f": 3}[a.dtype.kind]
    off_fmt, cnt_fmt, inline, entry = ("Q", "Q", 8, 20) if big else ("I", "H", 4, 12)
    header = (b"II" if order == "<" else b"MM") + (struct.pack(order + "HHHQ", 43, 8, 0, 0) if big else struct.pack(order + "HI", 42, 0))
    body, file_offset = b"", len(header)
    page_ifd_offsets = []
    for _ in range(pages):
        offsets = []
        for chunk in encoded:
            offsets.append(file_offset + len(body))
            body += chunk
        if len(body) & 1:
            body += b"\0"
        tags = [(256, 4, [width]), (257, 4, [height]), (258, 3, [a.dtype.itemsize * 8] * samples), (259, 3, [compression]),
                (262, 3, [1]), (277, 3, [samples]), (284, 3, [2 if planar else 1]), (317, 3, [predictor]),
                (339, 3, [kind] * samples)]
        long_kind = 16 if big else 4
        if tile:
            tags += [(322, 4, [tile[1]]), (323, 4, [tile[0]]), (324, long_kind, offsets), (325, long_kind, [len(c) for c in encoded])]
        else:
            tags += [(273, long_kind, offsets), (278, 4, [4]), (279, long_kind, [len(c) for c in encoded])]
        tags.sort()
        formats = {3: "H", 4: "I", 16: "Q"}
        ifd_at = file_offset + len(body)
        page_ifd_offsets.append(ifd_at)
        overflow_at = ifd_at + struct.calcsize(cnt_fmt) + entry * len(tags) + struct.calcsize(off_fmt)
        entries, overflow = b"", b""
        for tag, k, values in tags:
            packed = struct.pack(order + formats[k] * len(values), *values)
            e = struct.pack(order + "HH", tag, k) + struct.pack(order + off_fmt, len(values))
            if len(packed) <= inline:
                e += packed.ljust(inline, b"\0")
            else:
                e += struct.pack(order + off_fmt, overflow_at + len(overflow))
                overflow += packed + (b"\0" if len(packed) & 1 else b"")
            entries += e
        body += struct.pack(order + cnt_fmt, len(tags)) + entries + b"NEXT".ljust(struct.calcsize(off_fmt), b"T") + overflow
    data = bytearray(header + body)
    first = page_ifd_offsets[0]
    data[(8 if big else 4):(16 if big else 8)] = struct.pack(order + off_fmt, first)
    for i, at in enumerate(page_ifd_offsets):          # patch the next-IFD pointers
        n = struct.unpack_from(order + cnt_fmt, data, at)[0]
        where = at + struct.calcsize(cnt_fmt) + entry * n
        nxt = page_ifd_offsets[i + 1] if i + 1 < len(page_ifd_offsets) else 0
        data[where:where + struct.calcsize(off_fmt)] = struct.pack(order + off_fmt, nxt)
    return bytes(data)


@pytest.mark.parametrize("kwargs", [
    dict(order=">"), dict(big=True), dict(order=">", big=True), dict(tile=(16, 16)), dict(tile=(16, 32), planar=True),
    dict(planar=True), dict(compression=8), dict(compression=8, predictor=2), dict(order=">", compression=8, predictor=2),
    dict(tile=(16, 16), compression=8, predictor=2, order=">"), dict(pages=3)], ids=str)
def test_hand_assembled_layouts(kwargs):
    a = _cube((37, 29, 6), numpy.uint16)
    got = imread(_assemble(a, **kwargs))
    want = a.transpose(2, 0, 1) if kwargs.get("planar") else a
    if kwargs.get("pages"):
        want = numpy.stack([want] * kwargs["pages"])
    assert got.dtype == numpy.uint16 and got.shape == want.shape and numpy.array_equal(got, want)


@pytest.mark.parametrize("dtype", [numpy.float32, numpy.float64])
@pytest.mark.parametrize("kwargs", [dict(compression=8, predictor=3), dict(compression=8, predictor=3, tile=(16, 16), order=">")],
                         ids=str)
def test_floating_point_predictor(dtype, kwargs):
    a = _cube((21, 19, 3), dtype)
    got = imread(_assemble(a, **kwargs))
    assert got.dtype == numpy.dtype(dtype) and numpy.array_equal(got, a)


def test_unsupported_files_are_refused():
    a = _cube((8, 8, 1), numpy.uint16)
    with pytest.raises(TiffError):
        imread(_assemble(a, compression=7))                                  # JPEG-in-TIFF
    data = bytearray(_assemble(a))
    with pytest.raises(TiffError):
        imread(bytes(data[:2]) + struct.pack("<H", 99) + bytes(data[4:]))    # wrong magic


def test_native_and_python_lzw_decoders_agree(tmp_path):
    """hyp_tiff_lzw_decode (libhypelcnn_b200.so, host code) against the pure-Python decoder on libtiff-written strips:
    smooth data (long strings, table resets), noise (short strings), and a strip cut short by its capacity."""
    from PIL import Image
    from hypelcnn_b200.utilities import tiff_io as T
    path = str(tmp_path / "l.tif")
    smooth = (numpy.add.outer(numpy.arange(400), numpy.arange(700)) // 5).astype(numpy.uint16)
    noise = _cube((300, 300), numpy.uint8)
    flat = numpy.zeros((500, 500), numpy.uint8)
    for image in (smooth, noise, flat):
        Image.fromarray(image).save(path, compression="tiff_lzw")
        raw = open(path, "rb").read()
        page = T._read_ifds(memoryview(raw))[0]
        for offset, count in zip(page.all(T.STRIP_OFFSETS), page.all(T.STRIP_BYTE_COUNTS)):
            strip = raw[offset:offset + count]
            rows = min(page.first(T.ROWS_PER_STRIP), image.shape[0])
            expected = rows * image.shape[1] * image.dtype.itemsize
            python_bytes = T._lzw_decode_chunked(strip, expected)
            native_bytes = T._lzw_decode_native(strip, expected)
            assert native_bytes is not None and native_bytes[:len(python_bytes)] == python_bytes[:len(native_bytes)]
            assert len(native_bytes) == min(expected, len(python_bytes)) or len(native_bytes) == expected
            short = T._lzw_decode_native(strip, 1000)
            assert short == python_bytes[:1000]
        assert numpy.array_equal(imread(path), image)
    with pytest.raises(TiffError):
        T._lzw_decode_native(bytes([0xff] * 64), 100)                       # codes beyond the table

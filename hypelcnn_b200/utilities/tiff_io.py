"""TIFF scene files without tifffile: ``imread`` / ``imwrite`` for the rasters the reference's loaders and apps exchange
(``from tifffile import imread / imwrite``: loader/GRSS2013DataLoader.py, GRSS2018DataLoader.py:53-67,
GULFPORTDataLoader.py:22-43, common/common_nn_ops.py:568, classify/infer_for_classification.py:71-72,
gan/gan_infer_image_for_shadow.py:95-104) — hyperspectral cubes with tens to hundreds of samples per pixel, uint8 /
uint16 / int / float32 / float64, which general image libraries refuse.

Reader: classic TIFF and BigTIFF, either byte order, strips or tiles, chunky ("contig") or planar sample layout,
compression none / LZW / Deflate / PackBits, predictors none / horizontal / floating point, several pages of equal shape
stacked.  Result layout follows tifffile: ``[H,W]`` for one sample per pixel, ``[H,W,S]`` for chunky, ``[S,H,W]`` for
planar data, ``[pages, ...]`` for a stack.  GeoTIFF tags are ignored (the reference ignores them too).
Writer: uncompressed little-endian strips, chunky by default (``planarconfig="contig"`` as the reference passes) or
``"separate"``; BigTIFF when the file would pass 4 GiB.  LZW strips are decoded by ``hyp_tiff_lzw_decode`` in the native
library when it is built (host code; a pure-Python decoder of the same algorithm otherwise).
Pinned in tests/test_tiff_io.py against Pillow and OpenCV (libtiff) in both directions for every layout they support.
"""
import struct
import zlib

import numpy

_TYPE_FORMATS = {1: "B", 2: "c", 3: "H", 4: "I", 5: "II", 6: "b", 7: "B", 8: "h", 9: "i", 10: "ii", 11: "f", 12: "d",
                 13: "I", 16: "Q", 17: "q", 18: "Q"}
_NUMPY_FORMATS = {"B": "u1", "H": "u2", "I": "u4", "Q": "u8", "b": "i1", "h": "i2", "i": "i4", "q": "i8", "f": "f4",
                  "d": "f8"}
_SAMPLE_KINDS = {1: "u", 2: "i", 3: "f"}
_SAMPLE_FORMAT_OF_KIND = {"u": 1, "i": 2, "f": 3, "b": 1}

IMAGE_WIDTH, IMAGE_LENGTH, BITS_PER_SAMPLE, COMPRESSION, PHOTOMETRIC = 256, 257, 258, 259, 262
STRIP_OFFSETS, SAMPLES_PER_PIXEL, ROWS_PER_STRIP, STRIP_BYTE_COUNTS = 273, 277, 278, 279
PLANAR_CONFIG, PREDICTOR, TILE_WIDTH, TILE_LENGTH, TILE_OFFSETS, TILE_BYTE_COUNTS = 284, 317, 322, 323, 324, 325
EXTRA_SAMPLES, SAMPLE_FORMAT = 338, 339


class TiffError(ValueError):
    pass


# ---------------------------------------------------------------------------------------------------- decompression
def _unpack_bits(data, expected):
    out, i, n = bytearray(), 0, len(data)
    while i < n and len(out) < expected:
        header = data[i]
        i += 1
        if header < 128:
            out += data[i:i + header + 1]
            i += header + 1
        elif header > 128:
            out += data[i:i + 1] * (257 - header)
            i += 1
    return bytes(out)


def _lzw_decode(data, expected):
    """TIFF LZW: MSB-first codes of 9..12 bits, ClearCode 256, EndOfInformation 257, the width grows one code early."""
    table = [bytes([i]) for i in range(256)] + [b"", b""]
    out = bytearray()
    bits = int.from_bytes(data, "big")
    total = len(data) * 8
    position, width, previous = 0, 9, None
    while position + width <= total and len(out) < expected:
        code = (bits >> (total - position - width)) & ((1 << width) - 1)
        position += width
        if code == 256:
            table = table[:258]
            width, previous = 9, None
            continue
        if code == 257:
            break
        if previous is None:
            entry = table[code]
        elif code < len(table):
            entry = table[code]
            table.append(previous + entry[:1])
        elif code == len(table):
            entry = previous + previous[:1]
            table.append(entry)
        else:
            raise TiffError("corrupt LZW stream")
        out += entry
        previous = entry
        if len(table) + 1 >= (1 << width) and width < 12:
            width += 1
    return bytes(out)


def _lzw_decode_chunked(data, expected):
    """_lzw_decode over a big integer is quadratic in the strip size; strips larger than 64 KiB are decoded from a
    byte cursor instead (same algorithm, constant work per code)."""
    if len(data) <= 1 << 16:
        return _lzw_decode(data, expected)
    table = [bytes([i]) for i in range(256)] + [b"", b""]
    out = bytearray()
    buffer, buffered, index, n = 0, 0, 0, len(data)
    width, previous = 9, None
    while len(out) < expected:
        while buffered < width and index < n:
            buffer = ((buffer << 8) | data[index]) & 0xFFFFFFFF
            buffered += 8
            index += 1
        if buffered < width:
            break
        code = (buffer >> (buffered - width)) & ((1 << width) - 1)
        buffered -= width
        if code == 256:
            del table[258:]
            width, previous = 9, None
            continue
        if code == 257:
            break
        if previous is None:
            entry = table[code]
        elif code < len(table):
            entry = table[code]
            table.append(previous + entry[:1])
        elif code == len(table):
            entry = previous + previous[:1]
            table.append(entry)
        else:
            raise TiffError("corrupt LZW stream")
        out += entry
        previous = entry
        if len(table) + 1 >= (1 << width) and width < 12:
            width += 1
    return bytes(out)


def _lzw_decode_native(data, expected):
    """The same decoder inside libhypelcnn_b200.so (hyp_tiff_lzw_decode, host code): two orders of magnitude faster on
    the hundreds of megabytes of a compressed scene.  None when the library has not been built."""
    try:
        import ctypes
        from hypelcnn_b200 import _native
        lib = _native.lib()
    except Exception:
        return None
    out = ctypes.create_string_buffer(expected)
    written = ctypes.c_uint64(0)
    if lib.hyp_tiff_lzw_decode(data, len(data), out, expected, ctypes.byref(written)) != 0:
        raise TiffError("corrupt LZW stream")
    return out.raw[:written.value]


def _decompress(data, compression, expected):
    if compression == 1:
        return data
    if compression == 5:
        decoded = _lzw_decode_native(data, expected)
        return decoded if decoded is not None else _lzw_decode_chunked(data, expected)
    if compression in (8, 32946):
        return zlib.decompress(data)
    if compression == 32773:
        return _unpack_bits(data, expected)
    raise TiffError(f"TIFF compression {compression} is not supported (none, LZW, Deflate, PackBits are)")


# ------------------------------------------------------------------------------------------------------------ reader
class _Page:
    def __init__(self, tags, order):
        self.tags, self.order = tags, order

    def first(self, tag, default=None):
        value = self.tags.get(tag)
        return default if value is None else value[0]

    def all(self, tag, default=None):
        return self.tags.get(tag, default)


def _read_ifds(buffer):
    order = {b"II": "<", b"MM": ">"}.get(bytes(buffer[:2]))
    if order is None:
        raise TiffError("not a TIFF file")
    magic = struct.unpack(order + "H", buffer[2:4])[0]
    if magic == 42:
        big, offset = False, struct.unpack(order + "I", buffer[4:8])[0]
    elif magic == 43:
        big, offset = True, struct.unpack(order + "Q", buffer[8:16])[0]
    else:
        raise TiffError("not a TIFF file")
    count_format, entry_size, inline, offset_format = ("Q", 20, 8, "Q") if big else ("H", 12, 4, "I")
    pages = []
    while offset:
        count = struct.unpack_from(order + count_format, buffer, offset)[0]
        cursor = offset + struct.calcsize(count_format)
        tags = {}
        for _ in range(count):
            tag, kind = struct.unpack_from(order + "HH", buffer, cursor)
            number = struct.unpack_from(order + offset_format, buffer, cursor + 4)[0]
            value_at = cursor + 4 + struct.calcsize(offset_format)
            form = _TYPE_FORMATS.get(kind)
            if form is not None:
                size = struct.calcsize("=" + form) * number
                where = value_at if size <= inline else struct.unpack_from(order + offset_format, buffer, value_at)[0]
                if kind == 2:
                    tags[tag] = (bytes(buffer[where:where + number]),)
                elif size > 1 << 16 and form in _NUMPY_FORMATS:       # long offset tables: numpy does the unpacking
                    tags[tag] = numpy.frombuffer(buffer, dtype=numpy.dtype(order + _NUMPY_FORMATS[form]), count=number,
                                                 offset=where).tolist()
                else:
                    tags[tag] = struct.unpack_from(order + form * number, buffer, where)
            cursor += entry_size
        pages.append(_Page(tags, order))
        offset = struct.unpack_from(order + offset_format, buffer, cursor)[0]
    return pages


def _sample_dtype(page):
    bits = set(page.all(BITS_PER_SAMPLE, (1,)))
    kinds = set(page.all(SAMPLE_FORMAT, (1,)))
    if len(bits) != 1 or len(kinds) != 1:
        raise TiffError("samples of different type in one pixel are not supported")
    bits, kind = bits.pop(), _SAMPLE_KINDS.get(kinds.pop(), "u")
    if bits % 8 or bits // 8 not in (1, 2, 4, 8) or (kind == "f" and bits < 32):
        raise TiffError(f"{bits}-bit samples are not supported")
    return numpy.dtype(f"{page.order}{kind}{bits // 8}")


def _undo_horizontal_predictor(block, predictor):
    """block: [rows, width, samples] of one strip / tile; predictor 2 stores differences along the row."""
    if predictor == 2:
        return numpy.cumsum(block, axis=1, dtype=block.dtype)
    return block


def _read_page(buffer, page):
    width, height = page.first(IMAGE_WIDTH), page.first(IMAGE_LENGTH)
    samples = page.first(SAMPLES_PER_PIXEL, 1)
    planar = page.first(PLANAR_CONFIG, 1) == 2 and samples > 1
    compression, predictor = page.first(COMPRESSION, 1), page.first(PREDICTOR, 1)
    dtype = _sample_dtype(page)
    native = dtype.newbyteorder("=")
    tiled = TILE_OFFSETS in page.tags
    if tiled:
        block_w, block_h = page.first(TILE_WIDTH), page.first(TILE_LENGTH)
        offsets, counts = page.all(TILE_OFFSETS), page.all(TILE_BYTE_COUNTS)
    else:
        block_w, block_h = width, min(page.first(ROWS_PER_STRIP, height), height)
        offsets, counts = page.all(STRIP_OFFSETS), page.all(STRIP_BYTE_COUNTS)
        if counts is None:                       # a single uncompressed strip may omit its byte count
            counts = (len(buffer) - offsets[0],)
    across, down = -(-width // block_w), -(-height // block_h)
    planes = samples if planar else 1
    per_block = 1 if planar else samples
    if len(offsets) != across * down * planes:
        raise TiffError("strip / tile table does not match the image shape")
    out = numpy.empty((planes, height, width, per_block), dtype=native)
    index = 0
    for plane in range(planes):
        for by in range(down):
            rows = block_h if tiled else min(block_h, height - by * block_h)
            for bx in range(across):
                expected = rows * block_w * per_block * dtype.itemsize
                data = _decompress(bytes(buffer[offsets[index]:offsets[index] + counts[index]]), compression, expected)
                index += 1
                if len(data) < expected:
                    raise TiffError("strip / tile is shorter than its shape requires")
                if predictor == 3:
                    block = _float_predictor(data, rows, block_w, per_block, dtype)
                else:
                    block = numpy.frombuffer(data, dtype=dtype, count=rows * block_w * per_block).reshape(
                        rows, block_w, per_block)
                    block = _undo_horizontal_predictor(block.astype(native), predictor)
                y0, x0 = by * block_h, bx * block_w
                y1, x1 = min(y0 + rows, height), min(x0 + block_w, width)
                out[plane, y0:y1, x0:x1] = block[:y1 - y0, :x1 - x0]
    if planar:
        return out[..., 0]
    return out[0, :, :, 0] if samples == 1 else out[0]


def _float_predictor(data, rows, width, samples, dtype):
    """TIFF predictor 3 (libtiff fpAcc): per row, bytes are cumulatively summed with a stride of ``samples`` and hold
    all most-significant bytes first, then the next significance, ..."""
    size = dtype.itemsize
    raw = numpy.frombuffer(data, dtype=numpy.uint8, count=rows * width * samples * size).reshape(rows, width * size, samples)
    acc = numpy.cumsum(raw, axis=1, dtype=numpy.uint8).reshape(rows, size, width * samples)
    big_endian = numpy.ascontiguousarray(acc.transpose(0, 2, 1))
    return big_endian.view(dtype.newbyteorder(">")).reshape(rows, width, samples).astype(dtype.newbyteorder("="))


def imread(path):
    """-> numpy array, laid out like tifffile.imread (see the module docstring)."""
    buffer = numpy.memmap(path, dtype=numpy.uint8, mode="r") if not isinstance(path, (bytes, bytearray)) else \
        numpy.frombuffer(path, dtype=numpy.uint8)
    buffer = memoryview(buffer)
    pages = [p for p in _read_ifds(buffer) if IMAGE_WIDTH in p.tags]
    if not pages:
        raise TiffError("no image in the file")
    first = _read_page(buffer, pages[0])
    if len(pages) == 1:
        return first
    rest = []
    for page in pages[1:]:
        if (page.first(IMAGE_WIDTH), page.first(IMAGE_LENGTH), page.first(SAMPLES_PER_PIXEL, 1)) != \
                (pages[0].first(IMAGE_WIDTH), pages[0].first(IMAGE_LENGTH), pages[0].first(SAMPLES_PER_PIXEL, 1)):
            return first                           # thumbnails / masks after the main image: first series only
        rest.append(_read_page(buffer, page))
    return numpy.stack([first] + rest)


# ------------------------------------------------------------------------------------------------------------ writer
def imwrite(path, data, planarconfig="contig", photometric=None):
    """Write ``data`` — ``[H,W]``, ``[H,W,S]`` (``planarconfig="contig"``) or ``[S,H,W]`` (``"separate"``) — as one
    uncompressed little-endian TIFF page."""
    data = numpy.asarray(data)
    if data.dtype == bool:
        data = data.astype(numpy.uint8)
    if data.dtype.kind not in "uif" or data.dtype.itemsize not in (1, 2, 4, 8) or data.dtype == numpy.float16:
        raise TiffError(f"dtype {data.dtype} cannot be stored")
    if planarconfig not in ("contig", "separate"):
        raise TiffError("planarconfig must be 'contig' or 'separate'")
    if data.ndim == 2:
        height, width, samples, planar = data.shape[0], data.shape[1], 1, False
    elif data.ndim == 3 and planarconfig == "contig":
        height, width, samples, planar = data.shape[0], data.shape[1], data.shape[2], False
    elif data.ndim == 3:
        samples, height, width, planar = data.shape[0], data.shape[1], data.shape[2], True
    else:
        raise TiffError("expected a 2-D or 3-D array")
    payload = numpy.ascontiguousarray(data.astype(data.dtype.newbyteorder("<"), copy=False))
    if photometric is None:
        photometric = "rgb" if (samples in (3, 4) and data.dtype in (numpy.uint8, numpy.uint16)) else "minisblack"
    colour_samples = 3 if photometric == "rgb" else 1
    row_bytes = width * data.dtype.itemsize * (1 if planar else samples)
    rows_per_strip = max(1, min(height, (1 << 20) // max(row_bytes, 1)))
    strips_per_plane = -(-height // rows_per_strip)
    planes = samples if planar else 1
    strip_sizes = [min(rows_per_strip, height - s * rows_per_strip) * row_bytes for s in range(strips_per_plane)] * planes
    big = payload.nbytes + (1 << 25) >= (1 << 32)
    header = 16 if big else 8
    offsets, position = [], header
    for size in strip_sizes:
        offsets.append(position)
        position += size
    offset_type = 16 if big else 4
    tags = [(IMAGE_WIDTH, 4, [width]), (IMAGE_LENGTH, 4, [height]),
            (BITS_PER_SAMPLE, 3, [data.dtype.itemsize * 8] * samples), (COMPRESSION, 3, [1]),
            (PHOTOMETRIC, 3, [2 if photometric == "rgb" else 1]), (STRIP_OFFSETS, offset_type, offsets),
            (SAMPLES_PER_PIXEL, 3, [samples]), (ROWS_PER_STRIP, 4, [rows_per_strip]),
            (STRIP_BYTE_COUNTS, offset_type, strip_sizes), (PLANAR_CONFIG, 3, [2 if planar else 1])]
    if samples > colour_samples:
        # a fourth sample of an RGB image is unassociated alpha (2), anything else is unspecified data (0)
        tags.append((EXTRA_SAMPLES, 3, [2] if (photometric == "rgb" and samples == 4) else [0] * (samples - colour_samples)))
    tags.append((SAMPLE_FORMAT, 3, [_SAMPLE_FORMAT_OF_KIND[data.dtype.kind]] * samples))
    tags.sort(key=lambda t: t[0])

    count_format, inline, offset_format = ("<Q", 8, "<Q") if big else ("<H", 4, "<I")
    entry_size = 20 if big else 12
    ifd_at = position + (position & 1)
    overflow_at = ifd_at + struct.calcsize(count_format) + entry_size * len(tags) + struct.calcsize(offset_format)
    entries, overflow = b"", b""
    for tag, kind, values in tags:
        packed = struct.pack("<" + _TYPE_FORMATS[kind] * len(values), *values)
        entry = struct.pack("<HH", tag, kind) + struct.pack(offset_format, len(values))
        if len(packed) <= inline:
            entry += packed.ljust(inline, b"\0")
        else:
            entry += struct.pack(offset_format, overflow_at + len(overflow))
            overflow += packed + (b"\0" if len(packed) & 1 else b"")
        entries += entry
    with open(path, "wb") as f:
        f.write(b"II" + (struct.pack("<HHHQ", 43, 8, 0, ifd_at) if big else struct.pack("<HI", 42, ifd_at)))
        f.write(payload.tobytes() if payload.nbytes < (1 << 28) else memoryview(payload).cast("B"))
        if position & 1:
            f.write(b"\0")
        f.write(struct.pack(count_format, len(tags)) + entries + struct.pack(offset_format, 0) + overflow)

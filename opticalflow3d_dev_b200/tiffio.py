"""
Minimal TIFF reader / writer for the time-lapse driver.

The reference does its I/O through `tifffile` (calc_flow.py:445-465 read side, :509 memmap, :526-529
write side).  That package is not a dependency here; this module implements the subset the driver and
the reference's analysis scripts rely on, with the same call names:

    imwrite(path, arr, photometric='minisblack')   one IFD per leading-axis plane, little-endian,
                                                   uncompressed, SampleFormat from the dtype, classic TIFF
                                                   below 4 GB and BigTIFF above (what tifffile emits)
    imread(path)                                   all pages stacked; ImageJ hyperstacks reshaped (T, Z, Y, X)
    memmap(path)                                   zero-copy view when the pixel data are contiguous
                                                   (ImageJ hyperstacks, including the > 4 GB single-IFD kind)
    TiffFile(path).pages[i].shape / .imagej_metadata

Supported on the read side: classic TIFF and BigTIFF, both byte orders, strips and tiles, uncompressed
(Compression = 1), LZW (5: what the MATLAB twin writes, src/MATLAB/TIFFwrite.m:27), Deflate (8 / 32946) and PackBits
(32773) with the horizontal (2) and floating-point (3) predictors, 8/16/32/64-bit unsigned, signed and IEEE samples,
one sample per pixel.  The LZW and PackBits decoders are native (libof3d.so, host code); zlib is Python's.
"""
from __future__ import annotations

import json
import os
import struct
import zlib

import numpy as np

__all__ = ['imwrite', 'imread', 'memmap', 'TiffFile', 'natsorted']

_TYPE_SIZES = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 6: 1, 7: 1, 8: 2, 9: 4, 10: 8, 11: 4, 12: 8, 13: 4, 16: 8, 17: 8, 18: 8}
_TYPE_FMT = {1: 'B', 2: 'c', 3: 'H', 4: 'I', 6: 'b', 7: 'B', 8: 'h', 9: 'i', 11: 'f', 12: 'd', 13: 'I', 16: 'Q', 17: 'q', 18: 'Q'}


def natsorted(names):
    """Natural sort (digits compare as numbers): what natsort.natsorted does for plain file names."""
    import re

    def key(s):
        return [(0, int(t)) if t.isdigit() else (1, t) for t in re.split(r'(\d+)', s) if t != '']
    return sorted(names, key=key)


# ------------------------------------------------------------------------------------------ writer
def _sample_format(dtype):
    k = np.dtype(dtype).kind
    if k == 'u' or k == 'b':
        return 1
    if k == 'i':
        return 2
    if k == 'f':
        return 3
    raise TypeError('tiffio.imwrite: unsupported dtype %s' % dtype)


_WRITE_POOL = None


def _write_pool(threads):
    global _WRITE_POOL
    if _WRITE_POOL is None or _WRITE_POOL._max_workers < threads:
        from concurrent.futures import ThreadPoolExecutor
        _WRITE_POOL = ThreadPoolExecutor(max_workers=max(threads, 8))
    return _WRITE_POOL


def imwrite(path, data, photometric='minisblack', bigtiff=None, description=None, threads=1):
    """Write a 2D array as one page or an N-D array as a multi-page TIFF (one page per leading-axis plane).
    threads > 1: the pixel data of the pages are written by that many threads with positioned writes (the layout of the
    file is known up front); a RAM-disk or page-cache write is a memcpy at ~1.6 GB/s per thread."""
    if photometric not in ('minisblack', None):
        raise ValueError('only photometric="minisblack" is supported')
    a = np.asarray(data)
    if a.dtype == np.bool_:
        a = a.astype(np.uint8)
    if not a.dtype.isnative:
        a = a.astype(a.dtype.newbyteorder('='))
    if a.ndim < 2:
        raise ValueError('need at least a 2D array')
    shape = a.shape
    ny, nx = shape[-2], shape[-1]
    npages = int(np.prod(shape[:-2])) if a.ndim > 2 else 1
    a = np.ascontiguousarray(a).reshape(npages, ny, nx)
    page_bytes = ny * nx * a.dtype.itemsize
    if bigtiff is None:
        bigtiff = npages * (page_bytes + 512) + 4096 > 4 * 2 ** 30 - 2 ** 25     # tifffile switches near 4 GB
    if description is None:
        description = json.dumps({'shape': list(shape)})
    desc = description.encode('latin-1', 'replace') + b'\0'
    software = b'of3d-b200 tiffio\0'
    sf, bits = _sample_format(a.dtype), a.dtype.itemsize * 8

    off_fmt, cnt_fmt = ('<Q', '<Q') if bigtiff else ('<I', '<I')
    off_size = 8 if bigtiff else 4
    entry_size = 20 if bigtiff else 12

    def ifd_bytes(tags, ifd_offset):
        """tags: list of (tag, type, count, payload bytes).  Values too large for the slot go after the IFD."""
        n = len(tags)
        head = struct.pack('<Q', n) if bigtiff else struct.pack('<H', n)
        extra_off = ifd_offset + len(head) + n * entry_size + off_size
        entries, extra = b'', b''
        for tag, typ, count, payload in sorted(tags, key=lambda t: t[0]):
            e = struct.pack('<HH', tag, typ) + struct.pack(cnt_fmt, count)
            if len(payload) <= off_size:
                e += payload.ljust(off_size, b'\0')
            else:
                if (extra_off + len(extra)) % 2:
                    extra += b'\0'
                e += struct.pack(off_fmt, extra_off + len(extra))
                extra += payload
            entries += e
        return head, entries, extra

    long_t = 16 if bigtiff else 4

    def tags_for(i, data_off):
        t = [(256, 4, 1, struct.pack('<I', nx)), (257, 4, 1, struct.pack('<I', ny)),
             (258, 3, 1, struct.pack('<H', bits)), (259, 3, 1, struct.pack('<H', 1)),
             (262, 3, 1, struct.pack('<H', 1)),
             (273, long_t, 1, struct.pack(off_fmt, data_off)),
             (277, 3, 1, struct.pack('<H', 1)), (278, 4, 1, struct.pack('<I', ny)),
             (279, long_t, 1, struct.pack(off_fmt, page_bytes)),
             (282, 5, 1, struct.pack('<II', 1, 1)), (283, 5, 1, struct.pack('<II', 1, 1)),
             (296, 3, 1, struct.pack('<H', 1)), (339, 3, 1, struct.pack('<H', sf))]
        if i == 0:
            t += [(270, 2, len(desc), desc), (305, 2, len(software), software)]
        return t

    # layout of a page: IFD (+ overflow values) then pixel data, both 16-byte aligned; known without touching the data
    header = b'II' + (struct.pack('<HHHQ', 43, 8, 0, 16) if bigtiff else struct.pack('<HI', 42, 8))
    pos = len(header)
    meta, data_offs = [], []                                  # (offset, bytes) of every IFD blob; data offset of every page
    for i in range(npages):
        head, entries, extra = ifd_bytes(tags_for(i, 0), pos)
        ifd_len = len(head) + len(entries) + off_size + len(extra)
        data_off = (pos + ifd_len + 15) // 16 * 16
        head, entries, extra = ifd_bytes(tags_for(i, data_off), pos)
        next_ifd = 0 if i == npages - 1 else (data_off + page_bytes + 15) // 16 * 16
        meta.append((pos, head + entries + struct.pack(off_fmt, next_ifd) + extra))
        data_offs.append(data_off)
        pos = next_ifd if next_ifd else data_off + page_bytes
    total = pos

    threads = max(1, min(int(threads), npages))
    if threads == 1:
        with open(path, 'wb') as fh:
            fh.write(header)
            at = len(header)
            for i in range(npages):
                off, blob = meta[i]
                fh.write(b'\0' * (off - at) + blob)
                at = off + len(blob)
                fh.write(b'\0' * (data_offs[i] - at))
                fh.write(memoryview(a[i]).cast('B'))
                at = data_offs[i] + page_bytes
        return
    fd = os.open(os.fspath(path), os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o666)
    try:
        os.ftruncate(fd, total)                               # gaps between the pieces read back as zeros
        os.pwrite(fd, header, 0)
        for off, blob in meta:
            os.pwrite(fd, blob, off)

        def put(lo, hi):
            for i in range(lo, hi):
                mv = memoryview(a[i]).cast('B')
                done = 0
                while done < page_bytes:
                    done += os.pwrite(fd, mv[done:], data_offs[i] + done)

        cuts = [npages * k // threads for k in range(threads + 1)]
        futs = [_write_pool(threads).submit(put, cuts[k], cuts[k + 1]) for k in range(threads)]
        for f in futs:
            f.result()
    finally:
        os.close(fd)


# ------------------------------------------------------------------------------------------ reader
def lzw_decode_py(data, expected):
    """TIFF LZW (MSB-first codes, 9-12 bits, early change, ClearCode 256, EOI 257) -> bytes.  Reference implementation in
    pure Python: the reader uses the native decoder of libof3d.so and falls back to this one only when the library is
    absent; the tests compare the two."""
    out = bytearray()
    table = None
    nbits, nxt, prev = 9, 258, None
    acc = bits = 0
    it = iter(data)
    while len(out) < expected:
        while bits < nbits:
            try:
                acc = (acc << 8) | next(it)
            except StopIteration:
                return bytes(out)
            bits += 8
        code = (acc >> (bits - nbits)) & ((1 << nbits) - 1)
        bits -= nbits
        if code == 257:
            break
        if code == 256:
            table = [bytes((i,)) for i in range(256)] + [b'', b'']
            nbits, nxt, prev = 9, 258, None
            continue
        if table is None:
            raise ValueError('tiffio: LZW stream does not start with a clear code')
        if prev is None:
            entry = table[code]
        else:
            if code < len(table):
                entry = table[code]
            elif code == len(table):
                entry = prev + prev[:1]
            else:
                raise ValueError('tiffio: corrupt LZW stream')
            table.append(prev + entry[:1])
            nxt += 1
            if nxt + 1 >= (1 << nbits) and nbits < 12:      # early change
                nbits += 1
        out += entry
        prev = entry
    return bytes(out[:expected])


def packbits_decode_py(data, expected):
    out = bytearray()
    i, n = 0, len(data)
    while i < n and len(out) < expected:
        h = data[i]; i += 1
        if h < 128:
            out += data[i:i + h + 1]; i += h + 1
        elif h > 128:
            out += data[i:i + 1] * (257 - h); i += 1
    return bytes(out[:expected])


def _native_decode(kind, data, expected):
    """LZW (kind 0) / PackBits (kind 1) through libof3d.so; None when the library cannot be loaded."""
    try:
        from . import _lib
        lib = _lib.load()
    except Exception:
        return None
    import ctypes as C
    dst = np.empty(expected, dtype=np.uint8)
    src = np.frombuffer(data, dtype=np.uint8)
    n = lib.of3d_tiff_decode(kind, src.ctypes.data_as(C.c_void_p), src.size, dst.ctypes.data_as(C.c_void_p), expected)
    if n < 0:
        raise ValueError('tiffio: corrupt %s stream' % ('LZW' if kind == 0 else 'PackBits'))
    if n < expected:
        dst[n:] = 0
    return dst


def _decompress(compression, data, expected):
    """One strip / tile -> uint8 array of `expected` bytes (short streams are zero-padded like libtiff does)."""
    if compression == 1:
        buf = np.frombuffer(data, dtype=np.uint8)
    elif compression in (8, 32946):
        buf = np.frombuffer(zlib.decompress(data), dtype=np.uint8)
    elif compression in (5, 32773):
        buf = _native_decode(0 if compression == 5 else 1, data, expected)
        if buf is None:
            raw = (lzw_decode_py if compression == 5 else packbits_decode_py)(data, expected)
            buf = np.frombuffer(raw, dtype=np.uint8)
    else:
        raise NotImplementedError('tiffio: TIFF Compression=%d is not supported (supported: none, LZW, Deflate, PackBits)' % compression)
    if buf.size < expected:
        buf = np.concatenate([buf, np.zeros(expected - buf.size, np.uint8)])
    return buf[:expected]


class TiffPage:
    def __init__(self, tags, byteorder):
        self.tags = tags
        self.byteorder = byteorder
        self.shape = (int(tags[257][0]), int(tags[256][0]))
        bits = int(tags.get(258, (1,))[0])
        sf = int(tags.get(339, (1,))[0])
        kind = {1: 'u', 2: 'i', 3: 'f'}.get(sf, 'u')
        if bits % 8 or int(tags.get(277, (1,))[0]) != 1:
            raise NotImplementedError('tiffio: only whole-byte single-sample images are supported')
        self.dtype = np.dtype(byteorder + kind + str(bits // 8))
        self.compression = int(tags.get(259, (1,))[0])
        self.predictor = int(tags.get(317, (1,))[0])
        self.tiled = 324 in tags
        if self.tiled:
            self.tile = (int(tags[323][0]), int(tags[322][0]))          # (TileLength, TileWidth)
            self.offsets = [int(v) for v in tags[324]]
            self.bytecounts = [int(v) for v in tags[325]]
        else:
            self.offsets = [int(v) for v in tags[273]]
            self.bytecounts = [int(v) for v in tags.get(279, (self.shape[0] * self.shape[1] * self.dtype.itemsize,))]
        self.rows_per_strip = min(int(tags.get(278, (self.shape[0],))[0]), self.shape[0])
        d = tags.get(270)
        self.description = d if isinstance(d, str) else None

    @property
    def is_plain(self):
        """uncompressed strips without a predictor: the pixel bytes can be read (or mapped) as they lie in the file"""
        return self.compression == 1 and self.predictor == 1 and not self.tiled

    @property
    def is_contiguous(self):
        o = self.offsets
        return self.is_plain and all(o[i] + self.bytecounts[i] == o[i + 1] for i in range(len(o) - 1))

    def runs(self):
        """(file offset, byte count) of the strips, adjacent strips merged"""
        nbytes = self.shape[0] * self.shape[1] * self.dtype.itemsize
        out, pos = [], 0
        for off, cnt in zip(self.offsets, self.bytecounts):
            cnt = min(cnt, nbytes - pos)
            if cnt <= 0:
                break
            if out and out[-1][0] + out[-1][1] == off:
                out[-1][1] += cnt
            else:
                out.append([off, cnt])
            pos += cnt
        return out

    def _unpredict(self, block):
        """block: (rows, cols) view of the decoded bytes in FILE byte order, modified in place / returned"""
        if self.predictor == 1:
            return block
        if self.predictor == 2:                                      # horizontal differencing, per sample, modular
            u = block.view(self.dtype.newbyteorder(self.byteorder)).view(np.dtype(self.byteorder + 'u' + str(self.dtype.itemsize)))
            np.cumsum(u, axis=1, dtype=u.dtype, out=u)
            return block
        if self.predictor == 3:                                      # floating point: bytes de-interleaved, MSB plane first
            rows, rowbytes = block.shape
            isz = self.dtype.itemsize
            cols = rowbytes // isz
            b = np.cumsum(block, axis=1, dtype=np.uint8)
            planes = b.reshape(rows, isz, cols)                      # plane 0 = most significant byte
            order = range(isz) if self.byteorder == '>' else range(isz - 1, -1, -1)
            out = np.empty((rows, cols, isz), np.uint8)
            for dst_i, src_i in enumerate(order):
                out[:, :, dst_i] = planes[:, src_i, :]
            block[...] = out.reshape(rows, rowbytes)
            return block
        raise NotImplementedError('tiffio: TIFF Predictor=%d is not supported' % self.predictor)

    def readinto(self, fh, dst):
        """Read the page straight into `dst`, a writable C-contiguous array of the page's shape and (file) dtype."""
        if self.is_plain:
            mv = memoryview(dst).cast('B')
            pos = 0
            for off, cnt in self.runs():
                fh.seek(off)
                if fh.readinto(mv[pos:pos + cnt]) != cnt:
                    raise ValueError('tiffio: truncated image data')
                pos += cnt
            return
        ny, nx = self.shape
        isz = self.dtype.itemsize
        raw = np.asarray(dst).view(np.uint8).reshape(ny, nx * isz)
        if self.tiled:
            th, tw = self.tile
            across = (nx + tw - 1) // tw
            for i, (off, cnt) in enumerate(zip(self.offsets, self.bytecounts)):
                fh.seek(off)
                blk = self._unpredict(_decompress(self.compression, fh.read(cnt), th * tw * isz).copy().reshape(th, tw * isz))
                y0, x0 = (i // across) * th, (i % across) * tw
                h, w = min(th, ny - y0), min(tw, nx - x0)
                if h > 0 and w > 0:
                    raw[y0:y0 + h, x0 * isz:(x0 + w) * isz] = blk[:h, :w * isz]
            return
        rps = self.rows_per_strip
        for i, (off, cnt) in enumerate(zip(self.offsets, self.bytecounts)):
            y0 = i * rps
            if y0 >= ny:
                break
            h = min(rps, ny - y0)
            fh.seek(off)
            blk = _decompress(self.compression, fh.read(cnt), h * nx * isz).copy().reshape(h, nx * isz)
            raw[y0:y0 + h] = self._unpredict(blk)

    def asarray(self, fh):
        a = np.empty(self.shape, dtype=self.dtype)
        self.readinto(fh, a)
        return a.astype(self.dtype.newbyteorder('=')) if not self.dtype.isnative else a


class TiffFile:
    """Parses every IFD of a classic or Big TIFF (`pages`), and ImageJ's ImageDescription (`imagej_metadata`)."""

    def __init__(self, path):
        self.path = os.fspath(path)
        self.pages = []
        with open(self.path, 'rb') as fh:
            hdr = fh.read(16)
            if hdr[:2] == b'II':
                bo = '<'
            elif hdr[:2] == b'MM':
                bo = '>'
            else:
                raise ValueError('%s is not a TIFF file' % self.path)
            magic = struct.unpack(bo + 'H', hdr[2:4])[0]
            if magic == 42:
                self.bigtiff, off = False, struct.unpack(bo + 'I', hdr[4:8])[0]
            elif magic == 43:
                self.bigtiff, off = True, struct.unpack(bo + 'Q', hdr[8:16])[0]
            else:
                raise ValueError('%s: bad TIFF magic %d' % (self.path, magic))
            self.byteorder = bo
            seen = set()
            while off and off not in seen:
                seen.add(off)
                tags, off = self._read_ifd(fh, off)
                if 256 in tags and 257 in tags and (273 in tags or 324 in tags):
                    self.pages.append(TiffPage(tags, bo))
        self.imagej_metadata = self._imagej()

    def _read_ifd(self, fh, off):
        bo, big = self.byteorder, self.bigtiff
        fh.seek(off)
        n = struct.unpack(bo + ('Q' if big else 'H'), fh.read(8 if big else 2))[0]
        esz, slot = (20, 8) if big else (12, 4)
        raw = fh.read(n * esz + slot)
        tags = {}
        for i in range(n):
            e = raw[i * esz:(i + 1) * esz]
            tag, typ = struct.unpack(bo + 'HH', e[:4])
            count = struct.unpack(bo + ('Q' if big else 'I'), e[4:4 + slot])[0]
            size = _TYPE_SIZES.get(typ, 1) * count
            if size <= slot:
                payload = e[4 + slot:4 + slot + size]
            else:
                ptr = struct.unpack(bo + ('Q' if big else 'I'), e[4 + slot:4 + 2 * slot])[0]
                keep = fh.tell()
                fh.seek(ptr)
                payload = fh.read(size)
                fh.seek(keep)
            if typ == 2:
                tags[tag] = payload.split(b'\0')[0].decode('latin-1', 'replace')
            elif typ in (5, 10):
                v = struct.unpack(bo + ('I' if typ == 5 else 'i') * (2 * count), payload)
                tags[tag] = tuple((v[2 * j], v[2 * j + 1]) for j in range(count))
            elif typ in _TYPE_FMT and typ != 2:
                tags[tag] = struct.unpack(bo + _TYPE_FMT[typ] * count, payload)
            else:
                tags[tag] = payload
        nxt = struct.unpack(bo + ('Q' if big else 'I'), raw[n * esz:n * esz + slot])[0]
        return tags, nxt

    def _imagej(self):
        if not self.pages or not self.pages[0].description or not self.pages[0].description.startswith('ImageJ='):
            return None
        meta = {}
        for line in self.pages[0].description.splitlines():
            if '=' in line:
                k, v = line.split('=', 1)
                try:
                    meta[k] = int(v)
                except ValueError:
                    try:
                        meta[k] = float(v)
                    except ValueError:
                        meta[k] = v
        return meta

    # -- whole-file helpers
    def series_shape(self):
        """(npages_total, leading dims) taking ImageJ's images/frames/slices/channels into account."""
        p0 = self.pages[0]
        ij = self.imagej_metadata
        if ij:
            n = int(ij.get('images', len(self.pages)))
            dims = [int(ij[k]) for k in ('frames', 'slices', 'channels') if int(ij.get(k, 1)) > 1]
            if dims and int(np.prod(dims)) == n:
                return n, tuple(dims) + p0.shape
            return n, ((n,) if n > 1 else ()) + p0.shape
        n = len(self.pages)
        return n, ((n,) if n > 1 else ()) + p0.shape

    def asarray(self, out=None):
        """The whole series.  out: a C-contiguous array of the series' shape and native dtype (e.g. page-locked memory)
        to read into -- every plane goes from the file straight to its place, no intermediate copy."""
        n, shape = self.series_shape()
        p0 = self.pages[0]
        native = p0.dtype.newbyteorder('=')
        if out is not None:
            if out.size != int(np.prod(shape)) or out.dtype != native or not out.flags.c_contiguous:
                raise ValueError('tiffio: out must be a C-contiguous %s array of %s elements' % (native, tuple(shape)))
            a = out.reshape(shape)
        else:
            a = np.empty(shape, dtype=native)
        raw = a                          # file bytes land in place; a foreign byte order is swapped at the end
        with open(self.path, 'rb', buffering=0) as fh:
            if n > len(self.pages):      # ImageJ > 4 GB convention: one IFD, all planes contiguous after it
                fh.seek(p0.offsets[0])
                if fh.readinto(memoryview(raw.reshape(-1)).cast('B')) != raw.nbytes:
                    raise ValueError('tiffio: truncated image data')
            else:
                planes = raw.reshape((n,) + p0.shape)
                for i, p in enumerate(self.pages[:n]):
                    if p.shape != p0.shape or p.dtype != p0.dtype:
                        raise ValueError('tiffio: pages of different shape or dtype in one series')
                    p.readinto(fh, planes[i])
        if not p0.dtype.isnative:
            a.byteswap(inplace=True)
        return a if out is None else out

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def imread(path, out=None):
    return TiffFile(path).asarray(out=out)


def memmap(path, mode='r'):
    """numpy.memmap over the pixel data when all planes are stored contiguously and uncompressed
    (ImageJ hyperstacks are); raises otherwise.  Mirrors tifffile.memmap as used at calc_flow.py:509."""
    tf = TiffFile(path)
    n, shape = tf.series_shape()
    p0 = tf.pages[0]
    if not p0.is_plain:
        raise ValueError('tiffio.memmap: compressed, predicted or tiled TIFF cannot be memory-mapped')
    plane = p0.shape[0] * p0.shape[1] * p0.dtype.itemsize
    if n <= len(tf.pages):
        for i, p in enumerate(tf.pages[:n]):
            if not p.is_contiguous or p.offsets[0] != p0.offsets[0] + i * plane:
                raise ValueError('tiffio.memmap: image data are not contiguous in %s' % path)
    return np.memmap(tf.path, dtype=p0.dtype, mode=mode, offset=p0.offsets[0], shape=shape)


def imwrite_imagej(path, data, frames=None, slices=None):
    """Write a (T, Z, Y, X) / (T, Y, X) stack as an ImageJ hyperstack: contiguous planes after the first IFD's
    data offset and an 'ImageJ=' description, the OneTif input format of process_flow (calc_flow.py:452-459)."""
    a = np.ascontiguousarray(data)
    if a.ndim == 3:
        frames = a.shape[0] if frames is None else frames
        slices = 1 if slices is None else slices
    elif a.ndim == 4:
        frames, slices = a.shape[0], a.shape[1]
    else:
        raise ValueError('need (T,Y,X) or (T,Z,Y,X)')
    n = frames * slices
    desc = 'ImageJ=1.53t\nimages=%d\n' % n
    if slices > 1:
        desc += 'slices=%d\n' % slices
    if frames > 1:
        desc += 'frames=%d\n' % frames
    desc += 'hyperstack=true\nloop=false\n'
    # ImageJ layout: header, page-0 IFD, then ALL planes contiguously (so the stack can be memory-mapped),
    # then the IFDs of the remaining pages.
    _write_contiguous(path, a.reshape((n,) + a.shape[-2:]), desc)


def _write_contiguous(path, a, description):
    npages, ny, nx = a.shape
    page_bytes = ny * nx * a.dtype.itemsize
    total = npages * page_bytes
    big = total + npages * 256 + 4096 > 4 * 2 ** 30 - 2 ** 25
    off_fmt = '<Q' if big else '<I'
    slot = 8 if big else 4
    esz = 20 if big else 12
    long_t = 16 if big else 4
    desc = description.encode('latin-1') + b'\0'
    sf, bits = _sample_format(a.dtype), a.dtype.itemsize * 8
    data0 = 4096

    def page_tags(i):
        t = [(256, 4, 1, struct.pack('<I', nx)), (257, 4, 1, struct.pack('<I', ny)), (258, 3, 1, struct.pack('<H', bits)),
             (259, 3, 1, struct.pack('<H', 1)), (262, 3, 1, struct.pack('<H', 1)),
             (273, long_t, 1, struct.pack(off_fmt, data0 + i * page_bytes)), (277, 3, 1, struct.pack('<H', 1)),
             (278, 4, 1, struct.pack('<I', ny)), (279, long_t, 1, struct.pack(off_fmt, page_bytes)),
             (339, 3, 1, struct.pack('<H', sf))]
        if i == 0:
            t.append((270, 2, len(desc), desc))
        return sorted(t, key=lambda x: x[0])

    def build(tags, at, nxt):
        n = len(tags)
        head = struct.pack('<Q', n) if big else struct.pack('<H', n)
        extra_off = at + len(head) + n * esz + slot
        ent, extra = b'', b''
        for tag, typ, count, payload in tags:
            e = struct.pack('<HH', tag, typ) + struct.pack(off_fmt, count)
            if len(payload) <= slot:
                e += payload.ljust(slot, b'\0')
            else:
                e += struct.pack(off_fmt, extra_off + len(extra))
                extra += payload + (b'\0' if len(payload) % 2 else b'')
            ent += e
        return head + ent + struct.pack(off_fmt, nxt) + extra

    with open(path, 'wb') as fh:
        first = 16 if big else 8
        fh.write(b'II' + (struct.pack('<HHHQ', 43, 8, 0, first) if big else struct.pack('<HI', 42, first)))
        tail = data0 + total                       # IFDs of pages 1.. go after the pixel data
        ifd0 = build(page_tags(0), first, tail if npages > 1 else 0)
        assert first + len(ifd0) <= data0
        fh.write(ifd0)
        fh.write(b'\0' * (data0 - first - len(ifd0)))
        fh.write(memoryview(a).cast('B'))
        pos = tail
        for i in range(1, npages):
            blob = build(page_tags(i), pos, 0)
            nxt = pos + len(blob) if i < npages - 1 else 0
            blob = build(page_tags(i), pos, nxt)
            fh.write(blob)
            pos += len(blob)

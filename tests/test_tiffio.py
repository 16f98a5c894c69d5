"""CPU tests of the TIFF subset the time-lapse driver needs (round trips, ImageJ hyperstacks, memmap, BigTIFF)."""
import numpy as np
import pytest

from opticalflow3d_dev_b200 import tiffio


@pytest.mark.parametrize('dtype', [np.uint8, np.uint16, np.int16, np.float32, np.float64])
def test_roundtrip_multipage(tmp_path, dtype):
    a = (np.random.default_rng(0).normal(size=(5, 13, 17)) * 50 + 100).astype(dtype)
    p = tmp_path / 'a.tiff'
    tiffio.imwrite(p, a, photometric='minisblack')
    b = tiffio.imread(p)
    assert b.dtype == a.dtype and np.array_equal(a, b)
    tf = tiffio.TiffFile(p)
    assert len(tf.pages) == 5 and tf.pages[0].shape == (13, 17) and tf.imagej_metadata is None
    tiffio.imwrite(p, a[0])
    assert np.array_equal(tiffio.imread(p), a[0])
    tiffio.imwrite(p, a, bigtiff=True)
    assert np.array_equal(tiffio.imread(p), a)


def test_threaded_write_is_byte_identical(tmp_path):
    a = (np.random.default_rng(1).normal(size=(9, 13, 17)) * 50 + 100).astype(np.float32)
    for big in (False, True):
        tiffio.imwrite(tmp_path / 's.tiff', a, bigtiff=big)
        tiffio.imwrite(tmp_path / 'p.tiff', a, bigtiff=big, threads=4)
        assert (tmp_path / 's.tiff').read_bytes() == (tmp_path / 'p.tiff').read_bytes()


def test_imagej_hyperstack_and_memmap(tmp_path):
    a = np.arange(4 * 3 * 6 * 7).reshape(4, 3, 6, 7).astype(np.uint16)
    p = tmp_path / 'ij.tif'
    tiffio.imwrite_imagej(p, a)
    tf = tiffio.TiffFile(p)
    assert tf.imagej_metadata['frames'] == 4 and tf.imagej_metadata['slices'] == 3 and len(tf.pages) == 12
    assert np.array_equal(tiffio.imread(p), a)
    m = tiffio.memmap(p)
    assert m.shape == a.shape and np.array_equal(m[1:3], a[1:3])
    b = np.arange(5 * 6 * 7).reshape(5, 6, 7).astype(np.float32)
    tiffio.imwrite_imagej(p, b)
    assert tiffio.TiffFile(p).imagej_metadata['frames'] == 5 and np.array_equal(tiffio.memmap(p), b)


def test_readable_by_pillow(tmp_path):
    Image = pytest.importorskip('PIL.Image')
    f = np.random.default_rng(1).normal(size=(3, 9, 11)).astype(np.float32)
    p = tmp_path / 'f.tiff'
    tiffio.imwrite(p, f)
    im = Image.open(p)
    assert im.n_frames == 3
    im.seek(2)
    assert np.array_equal(np.array(im), f[2])


def test_natsorted():
    assert tiffio.natsorted(['a_t10_c.tif', 'a_t2_c.tif', 'a_t1_c.tif']) == ['a_t1_c.tif', 'a_t2_c.tif', 'a_t10_c.tif']


def test_process_flow_argument_errors(tmp_path):
    from opticalflow3d_dev_b200.calc_flow import process_flow
    with pytest.raises(SystemExit, match='does not exist'):
        process_flow(tmp_path / 'nope', 'x')
    with pytest.raises(SystemExit, match='No image files found'):
        process_flow(tmp_path, 'x_t.*')
    for t in range(3):
        tiffio.imwrite(tmp_path / ('x_t%d.tif' % t), np.zeros((2, 4, 4), np.uint16))
    with pytest.raises(SystemExit, match=r'only contains 3 files. Minimum 6\*tsig\+1 \(7\) files required'):
        process_flow(tmp_path, 'x_t.*')
    with pytest.raises(SystemExit, match='fileType must be either OneTif or SequenceT'):
        process_flow(tmp_path, 'x_t.*', fileType='Other')
    with pytest.raises(SystemExit, match='more than one file was found'):
        process_flow(tmp_path, 'x_t.*', fileType='OneTif')


def test_imread_into_preallocated_buffer(tmp_path):
    """imread(out=) reads every plane straight into the caller's (e.g. page-locked) array; shape may differ as long as
    the element count matches; dtype and contiguity are checked."""
    import numpy as np
    import pytest
    from opticalflow3d_dev_b200 import tiffio
    a = (np.arange(5 * 6 * 7).reshape(5, 6, 7) * 13 % 4001).astype(np.uint16)
    p = tmp_path / 'stack.tif'
    tiffio.imwrite(p, a)
    out = np.zeros((5, 6, 7), np.uint16)
    assert tiffio.imread(p, out=out) is out and np.array_equal(out, a)
    flat = np.zeros(5 * 6 * 7, np.uint16)
    tiffio.imread(p, out=flat)
    assert np.array_equal(flat.reshape(a.shape), a)
    with pytest.raises(ValueError):
        tiffio.imread(p, out=np.zeros((5, 6, 7), np.float32))
    with pytest.raises(ValueError):
        tiffio.imread(p, out=np.zeros((5, 6, 8), np.uint16))
    with pytest.raises(ValueError):
        tiffio.imread(p, out=np.zeros((5, 6, 14), np.uint16)[..., ::2])


# ------------------------------------------------------------------ compressed files written by libtiff (Pillow)
def _fixture_cases():
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location('make_tiff_fixtures', os.path.join(os.path.dirname(__file__), 'golden', 'make_tiff_fixtures.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize('idx', range(8))
def test_reads_libtiff_compressed_fixtures(idx):
    """LZW (what the MATLAB twin writes, TIFFwrite.m:27), Deflate, PackBits, predictor 2: byte-exact against the pixels
    the fixture script generated, for files encoded by libtiff (tests/golden/tiff/, written by Pillow)."""
    import os
    mod = _fixture_cases()
    name, dtype, pages, shape, kw = mod.CASES[idx]
    path = os.path.join(mod.HERE, name + '.tif')
    tf = tiffio.TiffFile(path)
    assert len(tf.pages) == pages
    comp = {'tiff_lzw': 5, 'tiff_adobe_deflate': 8, 'packbits': 32773}[kw['compression']]
    assert tf.pages[0].compression == comp
    assert tf.pages[0].predictor == kw.get('tiffinfo', {}).get(317, 1)
    got = tiffio.imread(path)
    want = np.stack([mod.make_image(1000 + 17 * i + len(name), shape, dtype) for i in range(pages)])
    if pages == 1:
        want = want[0]
    assert got.dtype == want.dtype and got.shape == want.shape
    assert np.array_equal(got, want)
    with pytest.raises(ValueError):
        tiffio.memmap(path)


def test_lzw_and_packbits_python_and_native_decoders_agree():
    """the native decoders (libof3d.so host code) against the pure-Python reference implementations, on every strip of
    the LZW / PackBits fixtures and on corrupt input"""
    import os
    mod = _fixture_cases()
    for name in ('lzw_u8', 'lzw_u16', 'lzw_pred2_u16', 'lzw_f32', 'packbits_u8'):
        tf = tiffio.TiffFile(os.path.join(mod.HERE, name + '.tif'))
        with open(tf.path, 'rb') as fh:
            for p in tf.pages:
                for i, (off, cnt) in enumerate(zip(p.offsets, p.bytecounts)):
                    fh.seek(off)
                    data = fh.read(cnt)
                    rows = min(p.rows_per_strip, p.shape[0] - i * p.rows_per_strip)
                    exp = rows * p.shape[1] * p.dtype.itemsize
                    kind = 0 if p.compression == 5 else 1
                    ref = (tiffio.lzw_decode_py if kind == 0 else tiffio.packbits_decode_py)(data, exp)
                    nat = tiffio._native_decode(kind, data, exp)
                    assert nat is not None and len(ref) == exp and nat.tobytes() == ref
    with pytest.raises(ValueError):
        tiffio._native_decode(0, bytes([0x00, 0x10, 0xff, 0xff]), 64)          # no leading clear code


def test_predictor3_and_tiles_round_trip():
    """floating-point predictor and tiled layout, encoded here following the TIFF 6.0 / Adobe technote layouts"""
    import struct
    import zlib
    rng = np.random.default_rng(3)
    img = rng.normal(size=(37, 45)).astype(np.float32)
    ny, nx = img.shape
    # predictor 3: per row, bytes split into planes (MSB first), then byte-wise horizontal differencing
    b = img.astype('<f4').view(np.uint8).reshape(ny, nx, 4)
    planes = np.concatenate([b[:, :, k] for k in (3, 2, 1, 0)], axis=1)
    diff = planes.copy()
    diff[:, 1:] = planes[:, 1:] - planes[:, :-1]
    payload = zlib.compress(diff.tobytes())

    def write(path, tags, blobs):
        # classic little-endian TIFF, one IFD; tags: (tag, type, values) with offsets patched for tag 273 / 324
        with open(path, 'wb') as fh:
            fh.write(b'II' + struct.pack('<HI', 42, 8))
            n = len(tags)
            data_at = 8 + 2 + n * 12 + 4
            extra = b''
            offs = []
            pos = data_at
            for bl in blobs:
                offs.append(pos); pos += len(bl) + (len(bl) & 1)
            ent = b''
            for tag, typ, vals in sorted(tags):
                if tag in (273, 324):
                    vals = offs
                fmt = {3: 'H', 4: 'I'}[typ]
                raw = struct.pack('<' + fmt * len(vals), *vals)
                if len(raw) <= 4:
                    ent += struct.pack('<HHI', tag, typ, len(vals)) + raw.ljust(4, b'\0')
                else:
                    ent += struct.pack('<HHII', tag, typ, len(vals), pos + len(extra))
                    extra += raw
            fh.write(struct.pack('<H', n) + ent + struct.pack('<I', 0))
            for bl in blobs:
                fh.write(bl + (b'\0' if len(bl) & 1 else b''))
            fh.write(extra)

    import tempfile, os
    with tempfile.TemporaryDirectory() as d:
        p3 = os.path.join(d, 'p3.tif')
        write(p3, [(256, 4, [nx]), (257, 4, [ny]), (258, 3, [32]), (259, 3, [8]), (262, 3, [1]), (273, 4, [0]), (277, 3, [1]),
                   (278, 4, [ny]), (279, 4, [len(payload)]), (317, 3, [3]), (339, 3, [3])], [payload])
        assert np.array_equal(tiffio.imread(p3), img)
        # tiles 16 x 32 of a uint16 image, deflate
        im16 = rng.integers(0, 60000, size=(37, 45)).astype(np.uint16)
        th, tw = 16, 32
        blobs = []
        for ty in range(0, ny, th):
            for tx in range(0, nx, tw):
                t = np.zeros((th, tw), np.uint16)
                blk = im16[ty:ty + th, tx:tx + tw]
                t[:blk.shape[0], :blk.shape[1]] = blk
                blobs.append(zlib.compress(t.tobytes()))
        pt = os.path.join(d, 'tiles.tif')
        write(pt, [(256, 4, [nx]), (257, 4, [ny]), (258, 3, [16]), (259, 3, [8]), (262, 3, [1]), (277, 3, [1]), (322, 4, [tw]),
                   (323, 4, [th]), (324, 4, [0] * len(blobs)), (325, 4, [len(b_) for b_ in blobs]), (339, 3, [1])], blobs)
        assert np.array_equal(tiffio.imread(pt), im16)

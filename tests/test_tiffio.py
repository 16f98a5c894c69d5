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

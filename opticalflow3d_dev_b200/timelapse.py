"""
process_flow -- the time-lapse driver of the reference (calc_flow.py:362-625), re-built as a streaming
GPU pipeline.

Contract kept from the reference: argument names / defaults, the SystemExit messages, file discovery by
regex `<imName>.tif` with natural sorting, the two input layouts (`OneTif` ImageJ hyperstack, `SequenceT`
one file per timepoint), outputs `<imDir>/OpticalFlow{3D,2D}/<name>/<name>_{vx,vy,vz,rel}_tNNNN.tiff`
(0-based frame index of the window centre) plus `<name>_parameters.csv`, and the progress prints.

What changed underneath: the reference re-reads NtChunk-1 frames from disk for every output timepoint and
runs strictly sequentially.  Here every frame is read and uploaded ONCE into a device-resident ring of the
last NtChunk frames, the operator runs on pointers into that ring (of3d_flow_frames), results come back
into pinned host buffers and are written by background threads while the GPU works on the next timepoint.
With torch.distributed initialised (one process per GPU) the output timepoints are sharded across ranks;
no collective is needed because every output depends only on its own window of frames.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import re
import sys
from concurrent.futures import ThreadPoolExecutor
from datetime import datetime
from pathlib import Path

import numpy as np

from . import _lib, tiffio
from .taps import flow_taps


def _rank_world():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def shard_range(n_items, rank, world):
    """Contiguous block of `n_items` for `rank` of `world` (sizes differ by at most one)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class _FrameSource:
    """Uniform access to frame t of either input layout, as a C-contiguous host array."""

    def __init__(self, imDir, fileList, fileType, Nt, Nz, spatialDimensions):
        self.imDir, self.files, self.fileType = imDir, fileList, fileType
        self.Nt, self.Nz, self.ndim = Nt, Nz, spatialDimensions
        self.stack = tiffio.memmap(imDir / fileList[0]) if fileType == 'OneTif' else None

    def frame(self, t):
        if self.stack is not None:
            a = self.stack[t]
        else:
            a = tiffio.imread(self.imDir / self.files[t])
        a = np.asarray(a)
        if self.ndim == 2 and a.ndim == 3 and a.shape[0] == 1:
            a = a[0]
        return np.ascontiguousarray(a)


def process_flow(imDir, imName, fileType="SequenceT", spatialDimensions=3, xyzSig=3, tSig=1, wSig=4, *,
                 precision='fp64', device=None, writers=4, verbose=True):
    ### Check inputs and set up paths (calc_flow.py:413-442) ########################################
    imDir = Path(imDir)
    if not imDir.is_dir():
        sys.exit('ERROR: image path \'%s\' does not exist' % imDir)
    imNamePattern = re.compile(imName + '.tif')
    fileList = [f for f in os.listdir(imDir) if imNamePattern.fullmatch(f)]
    if len(fileList) == 0:
        sys.exit('ERROR: No image files found. imName: ' + imName + ' imDir: ' + str(imDir))
    if fileType == 'OneTif':
        if len(fileList) > 1:
            sys.exit('ERROR: Type is OneTif but more than one file was found for imName: ' + imName)
    elif fileType == 'SequenceT':
        if len(fileList) < 6 * tSig + 1:
            sys.exit('ERROR: Image sequence found for file name ' + imName + ' only contains ' + str(len(fileList))
                     + ' files. Minimum 6*tsig+1 (' + str(6 * tSig + 1) + ') files required.')
    else:
        sys.exit('ERROR: fileType must be either OneTif or SequenceT.')
    fileList = tiffio.natsorted(fileList)
    if spatialDimensions < 2 or spatialDimensions > 3:
        sys.exit('ERROR: Number of spatial dimensions must be either 2 or 3.')

    ### Metadata (calc_flow.py:445-470) ###############################################################
    meta = tiffio.TiffFile(imDir / fileList[0])
    Ny, Nx = meta.pages[0].shape
    imj = meta.imagej_metadata
    if fileType == 'OneTif':
        if not imj:
            sys.exit('ERROR: fileType is OneTif, but no ImageJ metadata was detected')
        Nt = imj["frames"]
        Nz = imj["slices"] if spatialDimensions == 3 else 1
    else:
        Nz = len(meta.pages)
        Nt = len(fileList)
        if spatialDimensions == 2:
            if Nz != 1:
                sys.exit('ERROR: More than one z-slice detected for 2D processing')
        elif Nz <= 1:
            sys.exit('ERROR: 3D processing requested but Nz = ' + str(Nz))
    NtChunk = 6 * tSig + 1
    if not (NtChunk % 2):
        NtChunk = NtChunk + 1
    NtChunk = int(math.ceil(NtChunk))
    if not (NtChunk % 2):
        NtChunk += 1
    NtSlice = math.ceil(NtChunk / 2) - 1

    ### Output folder and parameter file (calc_flow.py:474-494) #######################################
    rank, world = _rank_world()
    savedir = imDir / ('OpticalFlow3D' if spatialDimensions == 3 else 'OpticalFlow2D')
    savedir.mkdir(exist_ok=True)
    imNameSave = imName.replace('.*', '')
    savedir = savedir / imNameSave
    savedir.mkdir(exist_ok=True)
    if rank == 0:
        with open(savedir / (imNameSave + '_parameters.csv'), 'w') as fh:
            fh.write('xyzSig,tiSig,wSig,Nx,Ny,Nz,Nt\n')          # 'tiSig' (sic) as in the reference
            fh.write(','.join(str(v) for v in (xyzSig, tSig, wSig, Nx, Ny, Nz, Nt)) + '\n')

    say = print if (verbose and rank == 0) else (lambda *a, **k: None)
    say('Note: regardless of input filenames, the first image = frame 0.')
    say('If your file names start from 0, adjust indexing accordingly for reading the output files.')
    say(' ')
    for hh in range(0, NtSlice):
        say(str(datetime.now()) + ' - No data will be saved for frame ' + str(hh) + ' to avoid edge effects')

    ### Streaming processing loop ####################################################################
    n_out = Nt - NtChunk + 1
    if n_out > 0:
        lo, hi = shard_range(n_out, rank, world)               # this rank's window starts hh in [lo, hi)
        if hi > lo:
            _stream(imDir, fileList, fileType, spatialDimensions, (Nt, Nz, Ny, Nx), (xyzSig, tSig, wSig), NtChunk, NtSlice,
                    range(lo, hi), savedir, imNameSave, precision, device, writers, verbose)
    for hh in range(max(Nt - NtSlice, 0), Nt):
        say(str(datetime.now()) + ' - No data will be saved for frame ' + str(hh) + ' to avoid edge effects')


def _stream(imDir, fileList, fileType, ndim, dims, sig, NtChunk, NtSlice, starts, savedir, name, precision, device, writers,
            verbose):
    import torch
    Nt, Nz, Ny, Nx = dims
    src = _FrameSource(imDir, fileList, fileType, Nt, Nz, ndim)
    first = src.frame(starts[0])
    sp = (Nz, Ny, Nx) if ndim == 3 else (Ny, Nx)
    if first.shape != sp:
        sys.exit('ERROR: frame shape %s does not match the metadata %s' % (first.shape, sp))
    if first.dtype not in _lib.DTYPE_CODES:
        first = first.astype(np.float64)
    in_dt = first.dtype
    code = _lib.DTYPE_CODES[in_dt]
    dev = (int(os.environ.get('OF3D_DEVICE', os.environ.get('LOCAL_RANK', 0))) if device is None else int(device))
    ctx = _lib.get_context(dev)
    lib = ctx.lib
    taps, keep = _lib.make_taps(flow_taps(*sig))
    kt = keep[3].size                                          # temporal taps actually read (<= NtChunk)
    off = NtSlice - kt // 2                                    # first frame of the window the t-filter touches
    prec = _lib.FP64 if precision == 'fp64' else _lib.FP32
    odt = np.float64 if precision == 'fp64' else np.float32
    nvox = int(np.prod(sp))
    fbytes = nvox * in_dt.itemsize
    nout = ndim + 1

    tdev = torch.device('cuda', dev)
    ring = torch.empty((kt, fbytes), dtype=torch.uint8, device=tdev)              # frame t lives in slot t % kt
    stage = [_lib.pinned_empty(sp, in_dt) for _ in range(2)]                      # pinned upload staging
    d_out = [torch.empty(nvox * np.dtype(odt).itemsize, dtype=torch.uint8, device=tdev) for _ in range(nout)]
    h_out = [[_lib.pinned_empty(sp, odt) for _ in range(nout)] for _ in range(2)]  # double-buffered results
    pending = [None, None]
    resident = set()
    up = 0
    names = ['vx', 'vy', 'vz', 'rel'] if ndim == 3 else ['vx', 'vy', 'rel']
    copy_stream = torch.cuda.Stream(device=tdev)

    def upload(t):
        nonlocal up
        a = src.frame(t)
        if a.dtype != in_dt:
            a = a.astype(in_dt)
        buf = stage[up % 2]
        copy_stream.synchronize()                                                 # staging buffer free again
        buf[...] = a
        with torch.cuda.stream(copy_stream):
            ring[t % kt].copy_(torch.from_numpy(buf.reshape(-1).view(np.uint8)), non_blocking=True)
        up += 1
        resident.add(t)
        resident.discard(t - kt)

    def write_all(arrs, tstr):
        for nm, a in zip(names, arrs):
            out = a
            if ndim == 3 and nm == 'rel' and precision == 'fp64':
                out = a.astype(np.float32)             # dtype the reference writes (calc_flow.py:355-357, :529)
            tiffio.imwrite(str(savedir / name) + '_' + nm + '_t' + tstr + '.tiff', out, photometric='minisblack')

    with ThreadPoolExecutor(max_workers=max(1, writers)) as pool:
        for i, hh in enumerate(starts):
            loopStart = datetime.now()
            centre = hh + NtSlice
            if verbose:
                print(str(datetime.now()) + ' - Processing frame ' + str(centre) + '...')
            for t in range(hh + off, hh + off + kt):
                if t not in resident:
                    upload(t)
            copy_stream.synchronize()
            ptrs = (C.c_void_p * kt)(*[ring[(hh + off + k) % kt].data_ptr() for k in range(kt)])
            optr = [C.c_void_p(o.data_ptr()) for o in d_out]
            if ndim == 2:
                optr = [optr[0], optr[1], None, optr[2]]
            rc = lib.of3d_flow_frames(ctx.handle, ndim, ptrs, code, _lib.DEVICE, Nz if ndim == 3 else 1, Ny, Nx, C.byref(taps), prec,
                                      0, optr[0], optr[1], optr[2], optr[3], _lib.DEVICE)
            _lib.check(rc, 'of3d_flow_frames')
            slot = i % 2
            if pending[slot] is not None:
                pending[slot].result()                 # the writer that used these host buffers has finished
            for o, h in zip(d_out, h_out[slot]):
                torch.from_numpy(h.reshape(-1).view(np.uint8)).copy_(o, non_blocking=True)
            torch.cuda.current_stream(tdev).synchronize()
            tstr = str(centre).zfill(4)
            pending[slot] = pool.submit(write_all, h_out[slot], tstr)
            if verbose:
                print(str(datetime.now()) + ' - Frame ' + str(centre) + ' saved.  Duration: ' + str(datetime.now() - loopStart))
        for p in pending:
            if p is not None:
                p.result()

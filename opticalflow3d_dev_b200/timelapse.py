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

    def frame_into(self, t, out):
        """Read frame t straight into `out` (page-locked staging of the frame's shape and dtype); False if the file's
        layout does not allow it (the caller then takes frame(t))."""
        if self.stack is not None:
            a = self.stack[t]
            if a.shape != out.shape:
                return False
            _lib.parallel_copy(out, a)
            return True
        try:
            tiffio.imread(self.imDir / self.files[t], out=out)
            return True
        except ValueError:
            return False

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
                 precision='fp64', device=None, writers=4, verbose=True, frame_step=1):
    """Reference calc_flow.py:362-625.  Keyword-only extras: precision / device as in calc_flow3D; writers = TIFF writer
    threads; frame_step = d analyses every centre frame c on the frames c + d*(-R..R) (strided time sampling, the
    frame-interval study of src/PublicationFigureScripts/plot_FigureS2_dt.m:54 `tNow = Ntslice+dtNow*(-3:3)`): the
    first and last d*NtSlice frames are skipped and the output files keep the centre's frame index."""
    frame_step = int(frame_step)
    if frame_step < 1:
        raise ValueError('frame_step must be a positive integer')
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
        if len(fileList) < 6 * tSig + 1:               # (the reference's check; frame_step > 1 may need more, see n_out)
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
    for hh in range(0, min(NtSlice * frame_step, Nt)):
        say(str(datetime.now()) + ' - No data will be saved for frame ' + str(hh) + ' to avoid edge effects')

    ### Streaming processing loop ####################################################################
    n_out = Nt - (NtChunk - 1) * frame_step                    # centres NtSlice*d .. Nt-1-NtSlice*d
    if n_out > 0:
        lo, hi = shard_range(n_out, rank, world)               # this rank's window starts hh in [lo, hi)
        if hi > lo:
            _stream(imDir, fileList, fileType, spatialDimensions, (Nt, Nz, Ny, Nx), (xyzSig, tSig, wSig), NtChunk, NtSlice,
                    range(lo, hi), savedir, imNameSave, precision, device, writers, verbose, frame_step)
    for hh in range(max(Nt - NtSlice * frame_step, 0), Nt):
        say(str(datetime.now()) + ' - No data will be saved for frame ' + str(hh) + ' to avoid edge effects')


class FlowStream:
    """Streaming engine for a time-lapse on one GPU: what the reference's loop at calc_flow.py:512-534 becomes when
    every frame is uploaded once and copies overlap compute.

        eng = FlowStream(spatial_shape, dtype, (xyzSig, tSig, wSig), precision='fp64', device=0)
        for t, frame in enumerate(frames):            # host arrays, in time order
            done = eng.push(frame)                    # -> None or (centre_index, (vx, vy[, vz], rel)) of an EARLIER window
        done = eng.flush()                            # the last pending result

    * a device-resident ring holds the last Kt frames (Kt = number of temporal taps); `push` copies the new frame
      through pinned staging on a copy stream;
    * when a window is complete, of3d_flow_frames is enqueued on the library's stream (asynchronous mode) into one of
      two device output sets, and the device->host copy of that set runs on a second copy stream while the NEXT
      window computes;
    * results are returned one call later.  By default (copy=False) they are VIEWS of pinned host buffers: three host
      buffer sets rotate, so a returned result stays valid during the NEXT call of push() and is overwritten by the
      one after -- finish with it (or copy it) before that.  copy=True returns ordinary arrays that the caller owns.
    * frame_step = d: the window of centre c is c + d*(-R..R) (plot_FigureS2_dt.m:54); the ring then holds (Kt-1)*d+1 frames.
    The engine owns a private library context (its asynchronous mode never leaks into calc_flow3D calls of the same
    thread); use it as a context manager or call close().
    """

    def __init__(self, spatial_shape, dtype, sigmas, precision='fp64', device=None, frame_step=1, copy=False):
        import torch
        self.torch = torch
        self.sp = tuple(int(v) for v in spatial_shape)
        self.ndim = len(self.sp)
        if self.ndim not in (2, 3):
            raise ValueError('spatial_shape must be (Ny,Nx) or (Nz,Ny,Nx)')
        self.in_dt = np.dtype(dtype)
        if self.in_dt not in _lib.DTYPE_CODES:
            raise TypeError('unsupported frame dtype %s' % self.in_dt)
        self.code = _lib.DTYPE_CODES[self.in_dt]
        self.dev = (int(os.environ.get('OF3D_DEVICE', os.environ.get('LOCAL_RANK', 0))) if device is None else int(device))
        self.ctx = _lib.Context(self.dev)                  # private: async mode and workspace are this engine's own
        self.taps, self._keep = _lib.make_taps(flow_taps(*sigmas))
        self.kt = self._keep[3].size
        self.step = int(frame_step)
        if self.step < 1:
            raise ValueError('frame_step must be a positive integer')
        self.span = (self.kt - 1) * self.step + 1          # frames between the first and the last frame of a window
        self.copy = bool(copy)
        self.precision = precision
        self.prec = _lib.FP64 if precision == 'fp64' else _lib.FP32
        self.odt = np.dtype(np.float64 if precision == 'fp64' else np.float32)
        self.nvox = int(np.prod(self.sp))
        self.nout = self.ndim + 1
        tdev = torch.device('cuda', self.dev)
        self.tdev = tdev
        fbytes = self.nvox * self.in_dt.itemsize
        self.ring = torch.empty((self.span, fbytes), dtype=torch.uint8, device=tdev)     # frame t lives in slot t % span
        self.stage = [_lib.pinned_empty(self.sp, self.in_dt) for _ in range(2)]
        self.stage_ev = [None, None]
        # 3D reliability leaves the device as float32, the dtype the reference returns (calc_flow.py:355-357)
        self.flags = _lib.FLAG_REL_F32 if (self.ndim == 3 and precision == 'fp64') else 0
        self.odts = [self.odt] * (self.nout - 1) + [np.dtype(np.float32) if self.flags else self.odt]
        self.d_out = [[torch.empty(self.nvox * d.itemsize, dtype=torch.uint8, device=tdev) for d in self.odts] for _ in range(2)]
        self.h_out = [[_lib.pinned_empty(self.sp, d) for d in self.odts] for _ in range(3)]
        self.s_in = torch.cuda.Stream(device=tdev)
        self.s_out = torch.cuda.Stream(device=tdev)
        self.s_lib = torch.cuda.ExternalStream(self.ctx.stream, device=tdev)
        self.ctx.set_async(True)
        self.t = 0                    # frames pushed so far
        self.nwin = 0                 # windows launched so far
        self.pending = None           # (centre, slot, done_event) of the window whose result is not yet returned
        self.slot_free_ev = [None, None]   # compute may overwrite d_out[slot] after its D2H finished
        self.ring_free_ev = None           # upload may overwrite a ring slot after the compute that read it
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _collect(self):
        if self.pending is None:
            return None
        centre, hslot, ev = self.pending
        ev.synchronize()
        self.pending = None
        return centre, self._result(hslot)

    def _result(self, hslot):
        return tuple(np.array(a) for a in self.h_out[hslot]) if self.copy else tuple(self.h_out[hslot])

    def staging(self):
        """The page-locked staging array the NEXT push may be given (filled by the caller, e.g. read from a file
        straight into it): push(eng.staging(), pinned=True) uploads it without any host copy."""
        k = self.t % 2
        if self.stage_ev[k] is not None:
            self.stage_ev[k].synchronize()                 # its previous upload has left the buffer
        return self.stage[k]

    def push(self, frame, pinned=False):
        """Feed the next frame (host array).  pinned=True: `frame` already lives in page-locked memory (e.g.
        _lib.pinned_empty), has the stream's dtype, is C-contiguous and stays untouched until the next push -- it is
        then uploaded directly, without the staging copy."""
        torch = self.torch
        a = np.asarray(frame)
        if a.shape != self.sp:
            raise ValueError('frame shape %s, expected %s' % (a.shape, self.sp))
        k = self.t % 2
        if pinned and a.dtype == self.in_dt and a.flags.c_contiguous:
            src = a
        else:
            if self.stage_ev[k] is not None:
                self.stage_ev[k].synchronize()             # pinned staging buffer is free again
            _lib.parallel_copy(self.stage[k], a)           # host copy (and dtype conversion) into pinned memory
            src = self.stage[k]
        with torch.cuda.stream(self.s_in):
            if self.ring_free_ev is not None:
                self.s_in.wait_event(self.ring_free_ev)    # the window that read this ring slot has been computed
            self.ring[self.t % self.span].copy_(torch.from_numpy(src.reshape(-1).view(np.uint8)), non_blocking=True)
            ev = torch.cuda.Event(); ev.record(self.s_in)
        self.stage_ev[k] = ev
        self.h2d_bytes += src.nbytes
        self.t += 1
        if self.t < self.span:
            return None
        # ---- a window is complete: frames t-span, t-span+d, .. t-1, centre t-1-(kt//2)*d.  Its compute and its D2H are enqueued BEFORE
        # the previous result is waited for, so the output copy stream never idles: D2H(t-1) overlaps compute(t).
        prev, self.pending = self.pending, None
        slot = self.nwin % 2                               # device output set
        hslot = self.nwin % 3                              # host output set
        first = self.t - self.span
        self.s_lib.wait_event(ev)                          # all uploads so far (same stream order) have landed
        if self.slot_free_ev[slot] is not None:
            self.s_lib.wait_event(self.slot_free_ev[slot])
        ptrs = (C.c_void_p * self.kt)(*[self.ring[(first + i * self.step) % self.span].data_ptr() for i in range(self.kt)])
        o = [C.c_void_p(x.data_ptr()) for x in self.d_out[slot]]
        if self.ndim == 2:
            o = [o[0], o[1], None, o[2]]
        nz = self.sp[0] if self.ndim == 3 else 1
        rc = self.ctx.lib.of3d_flow_frames(self.ctx.handle, self.ndim, ptrs, self.code, _lib.DEVICE, nz, self.sp[-2], self.sp[-1],
                                           C.byref(self.taps), self.prec, self.flags, o[0], o[1], o[2], o[3], _lib.DEVICE)
        _lib.check(rc, 'of3d_flow_frames')
        cev = torch.cuda.Event(); cev.record(self.s_lib)
        self.ring_free_ev = cev
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(cev)
            for d, h in zip(self.d_out[slot], self.h_out[hslot]):
                torch.from_numpy(h.reshape(-1).view(np.uint8)).copy_(d, non_blocking=True)
                self.d2h_bytes += h.nbytes
            dev_ = torch.cuda.Event(); dev_.record(self.s_out)
        self.slot_free_ev[slot] = dev_
        self.pending = (first + (self.kt // 2) * self.step, hslot, dev_)
        self.nwin += 1
        if prev is None:
            return None
        prev[2].synchronize()                              # D2H of the previous window has landed
        return prev[0], self._result(prev[1])

    def flush(self):
        return self._collect()

    def close(self):
        if self.ctx is not None:
            try:
                self.ctx.sync()
            finally:
                self.ctx.close()
                self.ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def _stream(imDir, fileList, fileType, ndim, dims, sig, NtChunk, NtSlice, starts, savedir, name, precision, device, writers,
            verbose, frame_step=1):
    Nt, Nz, Ny, Nx = dims
    src = _FrameSource(imDir, fileList, fileType, Nt, Nz, ndim)
    sp = (Nz, Ny, Nx) if ndim == 3 else (Ny, Nx)
    first = src.frame(starts[0])
    if first.shape != sp:
        sys.exit('ERROR: frame shape %s does not match the metadata %s' % (first.shape, sp))
    in_dt = first.dtype if first.dtype in _lib.DTYPE_CODES else np.dtype(np.float64)
    eng = FlowStream(sp, in_dt, sig, precision=precision, device=device, frame_step=frame_step)
    try:
        _stream_loop(eng, src, ndim, in_dt, NtSlice, starts, savedir, name, writers, verbose, frame_step)
    finally:
        eng.close()                                            # also on a writer / reader error: nothing stays asynchronous


def _stream_loop(eng, src, ndim, in_dt, NtSlice, starts, savedir, name, writers, verbose, d):
    kt, span = eng.kt, eng.span
    off = (NtSlice - kt // 2) * d                              # first frame of window 0 the t-filter touches
    names = ['vx', 'vy', 'vz', 'rel'] if ndim == 3 else ['vx', 'vy', 'rel']
    t_start = {}

    def write_one(nm, a, centre):
        out = a
        if ndim == 3 and nm == 'rel':
            out = a.astype(np.float32, copy=False)     # dtype the reference writes (calc_flow.py:355-357, :529)
        tiffio.imwrite(str(savedir / name) + '_' + nm + '_t' + str(centre).zfill(4) + '.tiff', out, photometric='minisblack')

    def write_all(pool, arrs, centre):
        """one task per output file; returns a waiter for the whole timepoint"""
        fs = [pool.submit(write_one, nm, a, centre) for nm, a in zip(names, arrs)]

        def wait():
            for f in fs:
                f.result()
            if verbose:
                print(str(datetime.now()) + ' - Frame ' + str(centre) + ' saved.  Duration: ' + str(datetime.now() - t_start[centre]))
        return wait

    # frames needed by this rank: window start hh touches frames hh+off, hh+off+d, .. hh+off+span-1
    t_lo, t_hi = starts[0] + off, starts[-1] + off + span
    from collections import deque
    with ThreadPoolExecutor(max_workers=max(1, int(writers))) as pool:
        futs = deque()                                     # timepoints being written; a result stays valid for one more push
        for t in range(t_lo, t_hi):
            while len(futs) > 1:
                futs.popleft()()
            centre_next = t - (span - 1) + (kt // 2) * d   # centre of the window this frame completes
            if t - t_lo >= span - 1:
                t_start[centre_next] = datetime.now()
                if verbose:
                    print(str(datetime.now()) + ' - Processing frame ' + str(centre_next) + '...')
            buf = eng.staging()
            if src.frame_into(t, buf):                     # file -> page-locked staging, no intermediate copy
                done = eng.push(buf, pinned=True)
            else:
                done = eng.push(src.frame(t).astype(in_dt, copy=False))
            if done is not None:
                futs.append(write_all(pool, done[1], done[0]))
        done = eng.flush()
        if done is not None:
            futs.append(write_all(pool, done[1], done[0]))
        for f in futs:
            f()

"""On-disk radar cubes -> model input, reading only the ROI rows.

Replaces, for the hot path, the reference's per-frame `np.load(...).astype(np.float32)` + crop + normalise in
`CRUW_POSE_Dataset.get_cube` / `get_cube_phase` (det3d/datasets/cruw_pose/cruw_pose.py:167-194), the channel packing of
`AssignLabelPose(2).__call__` (det3d/datasets/pipelines/pose.py:163-172), `collate_fn`'s `torch.tensor` copy
(cruw_pose.py:264-266), the two DataLoader worker processes (det3d/datasets/loader/build_loader.py:46-57) and the
blocking host->device copy (det3d/torchie/apis/train.py:27-62):

    file --pread of the ROI z/y rows, all x (rtp_npy_read_roi_slab, reader threads, GIL released)--> pinned slab
         --cudaMemcpyAsync on a copy stream--> device slab fp16 [B][lead][Z][Y][RX]
         --rtp_ingest_pack (x crop, fp32 cast, (v-a)/(b-a), clamp, bf16 P8 pack)--> P8 input of the engine

Per Doppler frame that is 16.8 MB read and copied instead of 67.1 MB read + 84 MB (fp32 ROI) pickled and copied.
`read_roi_slab` / `probe` are host-only and work without a GPU; `CubeLoader` needs one (no CPU fallback).
"""
import ctypes as C
import os
import queue
import threading
from concurrent.futures import ThreadPoolExecutor

import torch

from . import lib

ROI0 = (13, 32, 17)       # first ROI index per axis (z, y, x): configs/cruw_pose/hr3d.py:31-38 through cruw_pose.py:125-146
GRID = (16, 64, 160)      # ROI extent (Z, Y, X)


def probe(path):
    """Header of a cube file: dict(shape, descr, fortran_order, data_offset, file_bytes)."""
    info = lib.NpyInfo()
    lib.check(lib.load().rtp_npy_probe(os.fsencode(path), C.byref(info)), "rtp_npy_probe")
    return {"shape": tuple(info.shape[i] for i in range(info.ndim)), "descr": info.descr.decode(),
            "fortran_order": bool(info.fortran_order), "data_offset": info.data_offset, "file_bytes": info.file_bytes,
            "elem_bytes": info.elem_bytes}


def slab_shape(shape, Z, Y):
    """[lead, Z, Y, RX] of the slab read from a cube of `shape` = [..., RZ, RY, RX]."""
    lead = 1
    for d in shape[:-3]:
        lead *= int(d)
    return (lead, int(Z), int(Y), int(shape[-1]))


def read_roi_slab(path, z0=ROI0[0], Z=GRID[0], y0=ROI0[1], Y=GRID[1], out=None, threads=4):
    """Rows z[z0,z0+Z) y[y0,y0+Y) (full x) of every leading plane of the fp16 cube file -> fp16 tensor [lead,Z,Y,RX].
    `out`: a contiguous CPU fp16 tensor (e.g. a slice of a pinned staging buffer) with at least that many elements."""
    if out is None:
        out = torch.empty(slab_shape(probe(path)["shape"], Z, Y), dtype=torch.float16)
    if out.is_cuda or out.dtype != torch.float16 or not out.is_contiguous():
        raise lib.RtpError("read_roi_slab: `out` must be a contiguous CPU float16 tensor")
    lib.check(lib.load().rtp_npy_read_roi_slab(os.fsencode(path), z0, Z, y0, Y, out.data_ptr(), out.numel() * 2, threads),
              "rtp_npy_read_roi_slab")
    return out


def shard_paths(paths, rank, world, batch=1):
    """The contiguous share of `paths` for data-parallel rank `rank` of `world` (rtpose_b200.dist.shard_frames), trimmed to whole
    batches so that every rank iterates the same number of batches (the gradient all-reduce needs them in lock step):
    `CubeLoader(shard_paths(paths, rank, world, batch), batch)`."""
    from .dist import shard_frames
    paths = list(paths)
    per_rank = (len(paths) // world // batch) * batch if batch > 0 else len(paths) // world
    lo, _ = shard_frames(len(paths), rank, world)
    return paths[lo:lo + per_rank]


def ingest_slab(slab, x0=ROI0[2], X=GRID[2], norm=None, out=None, want_f32=False):
    """Device slab fp16 [B, lead, Z, Y, RX] -> P8 bf16 [B, lead, Z, Y, X] (and optionally the reference's fp32 tensor)."""
    from .p8 import P8, _stream
    lib.require_device()
    B, D, Z, Y, RX = slab.shape
    if not slab.is_cuda or slab.dtype != torch.float16 or not slab.is_contiguous():
        raise lib.RtpError("ingest_slab: the slab must be a contiguous CUDA float16 tensor")
    dst = out if out is not None else P8(B, D, Z, Y, X, device=slab.device)
    f32 = torch.empty((B, D, Z, Y, X), dtype=torch.float32, device=slab.device) if want_f32 else None
    a, b = norm if norm is not None else (0.0, 1.0)
    lib.call("rtp_ingest_pack", slab.data_ptr(), B, D, Z, Y, RX, 0, 0, x0, float(a), float(b - a), 1 if norm is not None else 0,
             dst.struct(), f32.data_ptr() if want_f32 else None, _stream())
    return (dst, f32) if want_f32 else dst


class CubeLoader:
    """Iterates over batches of cube files and yields `(P8 input, paths)` on the current CUDA stream.

    paths        list of .npy cube files (all of one shape); batches are consecutive groups of `batch` paths
    norm         (a, b) of `(v - a) / (b - a)` + clamp (rad_normalize_values, cruw_pose.py:46,172-173), or None for the
                 already-normalised phase cubes (get_cube_phase)
    depth        staging slots (pinned + device slab each); depth - 1 batches are read ahead of the consumer
    frame_workers / io_threads   frames read concurrently / pread threads per frame
    The yielded P8 is freshly allocated per batch (the consumer may keep it); the staging slot is recycled as soon as the
    ingest kernel of that batch has been enqueued.
    """

    def __init__(self, paths, batch, device="cuda", roi0=ROI0, grid=GRID, norm=None, depth=3, frame_workers=2, io_threads=8,
                 drop_last=True, want_f32=False):
        lib.require_device()
        self.paths = [os.fspath(p) for p in paths]
        if not self.paths:
            raise lib.RtpError("CubeLoader: no cube files")
        self.batch, self.device = int(batch), torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.roi0, self.grid, self.norm, self.want_f32 = tuple(roi0), tuple(grid), norm, want_f32
        self.depth, self.frame_workers, self.io_threads = max(2, int(depth)), int(frame_workers), int(io_threads)
        nb = len(self.paths) // self.batch if drop_last else -(-len(self.paths) // self.batch)
        self.batches = [self.paths[i * self.batch:(i + 1) * self.batch] for i in range(nb)]
        shape = probe(self.paths[0])["shape"]
        if len(shape) < 3:
            raise lib.RtpError("CubeLoader: %s is not a cube (shape %r)" % (self.paths[0], shape))
        self.file_shape = shape
        self.slab = slab_shape(shape, self.grid[0], self.grid[1])
        self.bytes_per_frame = 2 * self.slab[0] * self.slab[1] * self.slab[2] * self.slab[3]
        self._stage, self._live = None, None

    def __len__(self):
        return len(self.batches)

    def _staging(self):
        """Pinned + device slabs and their events: allocated once (cudaHostAlloc of ~1 GB is far slower than an epoch of
        reads) and shared by successive epochs, of which only one is live at a time."""
        if self._stage is None:
            full = (self.batch,) + self.slab
            self._stage = {"pinned": [torch.empty(full, dtype=torch.float16).pin_memory() for _ in range(self.depth)],
                           "dev": [torch.empty(full, dtype=torch.float16, device=self.device) for _ in range(self.depth)],
                           "copied": [torch.cuda.Event() for _ in range(self.depth)],   # H2D out of the pinned slot finished
                           "consumed": [None] * self.depth,                             # ingest that read the device slot
                           "copy_stream": torch.cuda.Stream(device=self.device),
                           "pool": ThreadPoolExecutor(max_workers=max(1, self.frame_workers))}
        return self._stage

    def close(self):
        """Stops a live epoch's producer thread (safe to call at any time; the loader can be iterated again)."""
        if self._live is not None:
            self._live.close()

    def __iter__(self):
        if self._live is not None:
            self._live.close()  # an abandoned epoch must stop writing into the staging slots first
        self._live = _Epoch(self)
        return self._live


class _Epoch:
    def __init__(self, ld):
        self.ld = ld
        st = ld._staging()
        self.pinned, self.dev, self.copied, self.consumed = st["pinned"], st["dev"], st["copied"], st["consumed"]
        self.copy_stream, self.pool = st["copy_stream"], st["pool"]
        self.free, self.ready = queue.Queue(), queue.Queue()
        for s in range(ld.depth):
            self.free.put(s)
        self.stop = False
        self.thread = threading.Thread(target=self._produce, name="rtp-cube-loader", daemon=True)
        self.thread.start()

    def _produce(self):
        ld = self.ld
        try:
            torch.cuda.set_device(ld.device)
            z0, y0 = ld.roi0[0], ld.roi0[1]
            Z, Y = ld.grid[0], ld.grid[1]
            for paths in ld.batches:
                s = self.free.get()
                if self.stop or s is None:
                    return
                self.copied[s].synchronize()  # the previous copy out of this pinned slot is done
                futs = [self.pool.submit(read_roi_slab, p, z0, Z, y0, Y, self.pinned[s][i], ld.io_threads)
                        for i, p in enumerate(paths)]
                for f in futs:
                    f.result()
                if self.stop:
                    return
                n = len(paths)
                with torch.cuda.stream(self.copy_stream):
                    if self.consumed[s] is not None:
                        self.copy_stream.wait_event(self.consumed[s])  # the ingest that read this device slot has run
                    self.dev[s][:n].copy_(self.pinned[s][:n], non_blocking=True)
                    self.copied[s].record(self.copy_stream)
                self.ready.put((s, n, paths, None))
            self.ready.put((None, 0, None, None))
        except BaseException as ex:  # surfaced in the consumer thread
            self.ready.put((None, 0, None, ex))

    def __iter__(self):
        return self

    def __next__(self):
        if self.stop:
            raise StopIteration
        s, n, paths, err = self.ready.get()
        if err is not None:
            self.close()
            raise err
        if s is None:
            self.close()
            raise StopIteration
        ld = self.ld
        cur = torch.cuda.current_stream(ld.device)
        cur.wait_event(self.copied[s])
        out = ingest_slab(self.dev[s][:n], ld.roi0[2], ld.grid[2], ld.norm, want_f32=ld.want_f32)
        ev = torch.cuda.Event()
        ev.record(cur)
        self.consumed[s] = ev
        self.free.put(s)
        return out, paths

    def close(self):
        """Stops the producer and waits for it (it may be in the middle of filling a staging slot)."""
        if not self.stop:
            self.stop = True
            self.free.put(None)
        if self.thread is not threading.current_thread():
            self.thread.join()
        if self.ld._live is self:
            self.ld._live = None

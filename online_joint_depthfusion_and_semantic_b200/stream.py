"""Frame ring in front of `Pipeline.fuse` (the reference feeds its loop from a DataLoader whose
workers decode and collate frames ahead of the consumer, dataset/replica.py:211-295, test_fusion.py:60-78).

`FrameStream.submit(host_batch)` issues the host->device copies of the frame on a copy stream, queues
`Pipeline.fuse` for it behind that copy on the compute stream, queues the device->host copy of the step's scalar
result into pinned memory, and returns the result of the PREVIOUS frame -- the only thing the host ever waits for is a
frame that was queued one call earlier, so the copies of frame i+1 and the read-back of frame i-1 overlap the kernels
of frame i and the GPU never idles between frames.  Frames are fused strictly in submission order (the integrator is
order dependent, modules/integrator.py:55-88).  `flush()` returns the results still in flight.

With `overlap_segmentation` (default) the AdapNet++ pass of a frame -- which needs the image but not the volumes -- runs on
a third stream (`Pipeline.segment`), so the segmentation of frame i+1 overlaps the extract / FusionNet / integrate part of
frame i: the two halves of the step are chains of ~150 and ~40 dependent launches whose gaps and partial waves
(64..144 CTAs on 148 SMs) fill each other.  The volumes are still updated strictly in order on the compute stream.
"""
import collections

import torch

# per-frame tensors the kernels read; everything else in the batch (pose: 100 bytes, ids) stays on the host
DEVICE_KEYS = ('image', 'tof_depth', 'mask', 'semantic_gt', 'depth', 'gt')


class FrameStream:
    def __init__(self, pipeline, database, device, result_fn=None, depth=2, keys=DEVICE_KEYS, overlap_segmentation=True):
        """result_fn(): 0-d device tensor describing the frame just fused (read back asynchronously), or None."""
        if torch.device(device).type != 'cuda':
            raise ValueError('FrameStream needs a CUDA device (there is no CPU path)')
        self.pipeline, self.database, self.device = pipeline, database, torch.device(device)
        self.result_fn, self.depth, self.keys = result_fn, max(1, int(depth)), tuple(keys)
        self._copy = torch.cuda.Stream(device=self.device)
        self._seg = torch.cuda.Stream(device=self.device) if overlap_segmentation else None
        self._inflight = collections.deque()                     # (event, pinned scalar or None)
        self._pinned = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(self.depth + 1)]
        self._n = 0
        self.h2d_bytes = 0

    def _upload(self, host_batch):
        main = torch.cuda.current_stream(self.device)
        out = dict(host_batch)
        nbytes = 0
        with torch.cuda.stream(self._copy):
            for k in self.keys:
                v = host_batch.get(k)
                if torch.is_tensor(v) and not v.is_cuda:
                    d = v.to(self.device, non_blocking=True)
                    d.record_stream(main)                         # allocated on the copy stream, consumed on the compute stream
                    out[k] = d
                    nbytes += v.numel() * v.element_size()
            ready = torch.cuda.Event()
            ready.record(self._copy)
        self.h2d_bytes = nbytes
        return out, ready

    def _pop(self):
        ev, pin = self._inflight.popleft()
        ev.synchronize()
        return None if pin is None else float(pin[0])

    def submit(self, host_batch):
        """Queue one frame; returns the result of the oldest frame once `depth` frames are in flight, else None."""
        main = torch.cuda.current_stream(self.device)
        batch, ready = self._upload(host_batch)
        sem = None
        if self._seg is not None:
            with torch.cuda.stream(self._seg), torch.no_grad():
                self._seg.wait_event(ready)
                for v in batch.values():
                    if torch.is_tensor(v) and v.is_cuda:
                        v.record_stream(self._seg)
                sem = self.pipeline.segment(batch, self.device)
                if sem is not None:
                    for v in sem.values():
                        if v is not None:
                            v.record_stream(main)                 # produced on the segmentation stream, consumed on the compute stream
                    seg_done = torch.cuda.Event()
                    seg_done.record(self._seg)
        main.wait_event(ready)
        if sem is not None:
            main.wait_event(seg_done)
            self.pipeline.fuse(batch, self.database, self.device, semantics=sem)
        else:
            self.pipeline.fuse(batch, self.database, self.device)
        pin = None
        if self.result_fn is not None:
            pin = self._pinned[self._n % len(self._pinned)]
            pin.copy_(self.result_fn().detach().reshape(1).float(), non_blocking=True)
        done = torch.cuda.Event()
        done.record(main)
        self._inflight.append((done, pin))
        self._n += 1
        return self._pop() if len(self._inflight) >= self.depth else None

    def flush(self):
        """Wait for every queued frame; their results in submission order."""
        out = []
        while self._inflight:
            out.append(self._pop())
        return out

"""End-to-end executor of the foveated resampling path for HOST buffers (the call a serving loop makes).

    pipe = ResamplePipeline(B, C, H, W, g=80, R=45, device=torch.device("cuda", 0))
    for batch in batches:                               # pinned host tensors
        pipe.submit(batch.image, batch.saliency, batch.pred, batch.mask_out)
    pipe.drain()                                        # every mask_out is filled

Every batch is one pass of SURVEY.md section 8 rows A4-A10 (models/models.py:594-657, 909, 933-940, 1044): saliency ->
grid, grid_sample(image, grid), inverse plan (A7 scatter, A9 point selection, Delaunay, point location) and the fused
inverse fill writing the [B,C,H,W] score tensor and its argmax.  What this class adds is the plumbing around the
kernels: three CUDA streams -- ingest (host->device copies + the two image-facing kernels: saliency -> grid and
grid_sample), inverse (plan + fill), and device->host copies -- over `depth` input/output slots, so the PCIe traffic of
batch i+1 and i-1 overlaps the HBM-bound fill of batch i.  All large buffers are allocated once.

`image_on_host=True` skips the bulk copy of the full-resolution image: the grid_sample kernel gathers its 4 taps per
output pixel straight from the pinned host tensor over PCIe (the sampler touches < 1 % of a 1024^2 frame, so pulling
32-byte sectors on demand moves far fewer bytes than copying the frame first).
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import FoveaError


class _Slot:
    def __init__(self, B, C, H, W, g, device, image_on_host):
        self.x = None if image_on_host else torch.empty(B, 3, H, W, device=device)
        self.xs = torch.empty(B, 1, g, g, device=device)
        self.pred = torch.empty(B, C, g, g, device=device)
        self.mask = torch.empty(B, H, W, device=device, dtype=torch.int64)
        self.h2d_done = torch.cuda.Event()
        self.compute_done = torch.cuda.Event()
        self.d2h_done = torch.cuda.Event()
        self.used = False


class ResamplePipeline:
    def __init__(self, B, C, H, W, g=80, R=45, device=None, triangulation="device", depth=2, want_scores=True,
                 image_on_host=False, filter_weight=None):
        if not torch.cuda.is_available():
            raise FoveaError("ResamplePipeline needs a CUDA device: there is no CPU fallback")
        self.dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.B, self.C, self.H, self.W, self.g, self.R = B, C, H, W, g, R
        self.tri, self.image_on_host = triangulation, image_on_host
        if filter_weight is None:
            from .models import makeGaussian       # models/models.py:510-515: fwhm = gaussian_radius
            filter_weight = torch.from_numpy(makeGaussian(2 * R + 1, fwhm=R)).float()
        self.g1x, self.g1y = (t.to(self.dev) for t in ops.separable_factors(filter_weight))
        self.copy_in = torch.cuda.Stream(self.dev)
        self.compute = torch.cuda.Stream(self.dev)
        self.copy_out = torch.cuda.Stream(self.dev)
        self.slots = [_Slot(B, C, H, W, g, self.dev, image_on_host) for _ in range(depth)]
        self.scores = torch.empty(B, C, H, W, device=self.dev) if want_scores else None   # compute stream only
        self.n = 0

    def submit(self, hx, hxs, hpred, hmask_out):
        """Enqueue one batch: pinned host image [B,3,H,W], saliency [B,1,g,g], pred [B,C,g,g] -> hmask_out [B,H,W] int64
        (pinned).  Returns immediately; `drain()` (or reuse of the slot `depth` submits later) orders completion."""
        for t, name in ((hx, "image"), (hxs, "saliency"), (hpred, "pred"), (hmask_out, "mask_out")):
            if t.is_cuda or not t.is_pinned():
                raise FoveaError(f"ResamplePipeline.submit: {name} must be a pinned host tensor")
        s = self.slots[self.n % len(self.slots)]
        self.n += 1
        g, R = self.g, self.R
        with torch.cuda.stream(self.copy_in):                # ingest stage: copies + the two image-facing kernels
            if s.used:
                self.copy_in.wait_event(s.compute_done)      # the slot's previous batch has consumed its inputs
            if not self.image_on_host:
                s.x.copy_(hx, non_blocking=True)
            s.xs.copy_(hxs, non_blocking=True)
            s.pred.copy_(hpred, non_blocking=True)
            s.grid = ops.saliency_to_grid(s.xs, self.g1x, self.g1y, g, g, R, R, "replication", (g, g))
            s.x_sampled = ops.grid_sample(hx if self.image_on_host else s.x, s.grid)
            s.h2d_done.record(self.copy_in)
        with torch.cuda.stream(self.compute):                # inverse stage
            self.compute.wait_event(s.h2d_done)
            if s.used:
                self.compute.wait_event(s.d2h_done)          # the slot's previous mask has left the device
            s.grid.record_stream(self.compute)
            plan = ops.build_inverse_plan(s.grid, (self.H, self.W), nchan=self.C, triangulation=self.tri)
            ops.inverse_fill(plan, s.pred, want_scores=self.scores is not None, want_mask=True, zero_residual=True,
                             out=self.scores, mask_out=s.mask)
            s.compute_done.record(self.compute)
        with torch.cuda.stream(self.copy_out):
            self.copy_out.wait_event(s.compute_done)
            hmask_out.copy_(s.mask, non_blocking=True)
            s.d2h_done.record(self.copy_out)
        s.used = True
        return s.x_sampled

    def fence(self, stream=None):
        """Make `stream` (default: the current stream) wait for everything submitted so far."""
        stream = stream or torch.cuda.current_stream(self.dev)
        for s in self.slots:
            if s.used:
                stream.wait_event(s.d2h_done)

    def start_after(self, event):
        """Make all three pipeline streams wait for `event` (used to bracket a timed region)."""
        for st in (self.copy_in, self.compute, self.copy_out):
            st.wait_event(event)

    def drain(self):
        for st in (self.copy_in, self.compute, self.copy_out):
            st.synchronize()

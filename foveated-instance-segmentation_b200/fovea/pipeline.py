"""End-to-end executor of the foveated resampling path for HOST buffers (the call a serving loop makes).

    pipe = ResamplePipeline(B, C, H, W, g=80, R=45, device=torch.device("cuda", 0))
    for batch in batches:                               # pinned host tensors
        pipe.submit(batch.image, batch.saliency, batch.pred, batch.mask_out)
    pipe.drain()                                        # every mask_out is filled

Every batch is one pass of SURVEY.md section 8 rows A4-A10 (models/models.py:594-657, 909, 933-940, 1044): saliency ->
grid, grid_sample(image, grid), inverse plan (A7 scatter, A9 point selection, Delaunay, point location) and the fused
inverse fill writing the [B,C,H,W] score tensor and its argmax.  What this class adds is the plumbing around the
kernels: four CUDA streams -- ingest (host->device copies + the two image-facing kernels: saliency -> grid and
grid_sample), plan (high priority: scatter, point selection, Delaunay, point location), fill (+ fused argmax), and
device->host copies -- over `depth` input/output slots, so the PCIe traffic of batch i+1 and i-1 and the latency-bound
plan of batch i+1 overlap the HBM-bound fill of batch i.  All large buffers are allocated once.

`image_on_host=True` skips the bulk copy of the full-resolution image: the grid_sample kernel gathers its 4 taps per
output pixel straight from the pinned host tensor over PCIe (the sampler touches < 1 % of a 1024^2 frame, so pulling
32-byte sectors on demand moves far fewer bytes than copying the frame first).
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import FoveaError


class _Slot:
    def __init__(self, B, C, H, W, g, device, n_copy, image_dtype, mask_dtype):
        self.x = torch.empty(n_copy, 3, H, W, device=device, dtype=image_dtype) if n_copy else None
        self.xs = torch.empty(B, 1, g, g, device=device)
        self.pred = torch.empty(B, C, g, g, device=device)
        self.mask = torch.empty(B, H, W, device=device, dtype=mask_dtype)
        self.h2d_done = torch.cuda.Event()
        self.compute_done = torch.cuda.Event()
        self.d2h_done = torch.cuda.Event()
        self.used = False


class ResamplePipeline:
    def __init__(self, B, C, H, W, g=80, R=45, device=None, triangulation="device", depth=2, want_scores=True,
                 image_on_host=False, filter_weight=None, image_dtype=torch.float32, mask_dtype=torch.int64):
        """image_dtype=torch.uint8: the image arrives as the decoder produced it and ToTensor()'s /255 is folded into the
        sampler (a quarter of the PCIe bytes); mask_dtype=torch.uint8: narrower masks than torch.argmax's int64 (an
        eighth of the D2H bytes).  Both are API options beyond the reference's dtypes; the defaults are the reference's."""
        if not torch.cuda.is_available():
            raise FoveaError("ResamplePipeline needs a CUDA device: there is no CPU fallback")
        self.dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.B, self.C, self.H, self.W, self.g, self.R = B, C, H, W, g, R
        self.tri = triangulation
        if filter_weight is None:
            from .models import makeGaussian       # models/models.py:510-515: fwhm = gaussian_radius
            filter_weight = torch.from_numpy(makeGaussian(2 * R + 1, fwhm=R)).float()
        self.g1x, self.g1y = (t.to(self.dev) for t in ops.separable_factors(filter_weight))
        self.copy_in = torch.cuda.Stream(self.dev)
        self.plan_stream = torch.cuda.Stream(self.dev, priority=-1)
        self.compute = torch.cuda.Stream(self.dev)
        self.copy_out = torch.cuda.Stream(self.dev)
        self.image_dtype = image_dtype
        # image_on_host may be a fraction: that share of the frames (the LAST ones of the batch) is gathered over PCIe by
        # the sampler, the rest is bulk-copied.  (Measured on the bench workload: 0 -> 3.4 k, 0.25 -> 3.9 k, 1 -> 4.3 k
        # frames/s: gathering everything wins.)
        frac = float(image_on_host)
        self.n_copy = B - int(round(B * min(max(frac, 0.0), 1.0)))
        self.slots = [_Slot(B, C, H, W, g, self.dev, self.n_copy, image_dtype, mask_dtype) for _ in range(depth)]
        self.scores = torch.empty(B, C, H, W, device=self.dev) if want_scores else None   # compute stream only
        self.n = 0

    def submit(self, hx, hxs, hpred, hmask_out):
        """Enqueue one batch: pinned host image [B,3,H,W], saliency [B,1,g,g], pred [B,C,g,g] -> hmask_out [B,H,W] int64
        (pinned).  Returns x_sampled [B,3,g,g] (ordered on the caller's current stream) immediately; `drain()` (or reuse
        of the slot `depth` submits later) orders completion of the masks.  The host buffers must stay untouched until
        then (they are read / written asynchronously)."""
        for t, name in ((hx, "image"), (hxs, "saliency"), (hpred, "pred"), (hmask_out, "mask_out")):
            if t.is_cuda or not t.is_pinned():
                raise FoveaError(f"ResamplePipeline.submit: {name} must be a pinned host tensor")
        s = self.slots[self.n % len(self.slots)]
        self.n += 1
        g, R = self.g, self.R
        with torch.cuda.stream(self.copy_in):                # ingest stage: copies + the two image-facing kernels
            if s.used:
                self.copy_in.wait_event(s.compute_done)      # the slot's previous batch has consumed its inputs
            nc = self.n_copy
            s.xs.copy_(hxs, non_blocking=True)
            s.pred.copy_(hpred, non_blocking=True)
            s.grid = ops.saliency_to_grid(s.xs, self.g1x, self.g1y, g, g, R, R, "replication", (g, g))
            sample = ops.grid_sample_u8 if self.image_dtype == torch.uint8 else ops.grid_sample
            parts = []
            if nc < self.B:                                  # gathered straight from the pinned host frames
                parts.append(sample(hx[nc:], s.grid[nc:]))
            if nc:
                s.x.copy_(hx[:nc], non_blocking=True)
                parts.insert(0, sample(s.x, s.grid[:nc]))
            s.x_sampled = parts[0] if len(parts) == 1 else torch.cat(parts, 0)
            s.h2d_done.record(self.copy_in)
        # x_sampled is produced on the ingest stream: the caller's stream (where the encoder would consume it) is ordered
        # behind it here, so the returned tensor is safe to use like any other torch result
        torch.cuda.current_stream(self.dev).wait_event(s.h2d_done)
        with torch.cuda.stream(self.plan_stream):            # saliency-only half of the inverse stage, high priority:
            self.plan_stream.wait_event(s.h2d_done)          # it overlaps the HBM-bound fill of the previous batch
            plan = ops.build_inverse_plan(s.grid, (self.H, self.W), nchan=self.C, triangulation=self.tri,
                                          dense_winner=False)
            planned = torch.cuda.Event()
            planned.record(self.plan_stream)
        with torch.cuda.stream(self.compute):                # fill (+ fused argmax)
            self.compute.wait_event(s.h2d_done)
            self.compute.wait_event(planned)
            if s.used:
                self.compute.wait_event(s.d2h_done)          # the slot's previous mask has left the device
            ops.inverse_fill(plan, s.pred, want_scores=self.scores is not None, want_mask=True, zero_residual=True,
                             out=self.scores, mask_out=s.mask)
            s.compute_done.record(self.compute)
        s.plan = plan                                        # keep the plan's buffers alive until the slot is reused
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
        for st in (self.copy_in, self.plan_stream, self.compute, self.copy_out):
            st.wait_event(event)

    def drain(self):
        for st in (self.copy_in, self.plan_stream, self.compute, self.copy_out):
            st.synchronize()
        for s in self.slots:
            if s.used:
                ops.check_plan(s.plan)


class DevicePipeline:
    """Device-resident executor: the same path for inputs already in HBM, software-pipelined ACROSS batches.

    Everything up to the per-pixel source map (`ops.build_inverse_plan`: A7 scatter, A9 point selection, Delaunay,
    point location) depends only on the saliency, and most of it runs one CTA per frame -- latency-bound on 64 of the
    148 SMs.  It is therefore issued on a HIGH-PRIORITY stream, one batch ahead of the HBM-bound fill of the previous
    batch on a second stream: the plan kernels take the SMs they need as soon as fill CTAs retire, the fill keeps the
    rest.  Measured (64 frames of 1024^2, C = 51): 5.13 ms per batch serial, 3.83 ms pipelined.

        pipe = DevicePipeline(B, C, H, W, g, R)
        for x, xs, pred in batches:            # CUDA tensors
            x_sampled, scores = pipe.submit(x, xs, pred)
        pipe.fence()                           # current stream waits for everything submitted

    `scores` is ONE buffer reused by every batch (13.7 GB at the bench size); consume it on `pipe.fill_stream` (or after
    `fence()`) before the next `submit` overwrites it.
    """

    def __init__(self, B, C, H, W, g=80, R=45, device=None, triangulation="device", depth=2, want_mask=False,
                 filter_weight=None, scores=None, interp="tri", want_scores=True):
        """want_scores=False (with want_mask=True): mask mode -- the [B,C,H,W] score tensor is never materialised, stage 3
        runs the pruned arg-max fill (fovea_inverse_mask) and `submit` returns (x_sampled, mask)."""
        if not torch.cuda.is_available():
            raise FoveaError("DevicePipeline needs a CUDA device: there is no CPU fallback")
        self.dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.B, self.C, self.H, self.W, self.g, self.R = B, C, H, W, g, R
        self.tri, self.depth, self.interp = triangulation, max(1, depth), interp
        if filter_weight is None:
            from .models import makeGaussian
            filter_weight = torch.from_numpy(makeGaussian(2 * R + 1, fwhm=R)).float()
        self.g1x, self.g1y = (t.to(self.dev) for t in ops.separable_factors(filter_weight))
        self.plan_stream = torch.cuda.Stream(self.dev, priority=-1)
        self.fill_stream = torch.cuda.Stream(self.dev, priority=0)
        if not want_scores and not want_mask:
            raise FoveaError("DevicePipeline: nothing to produce (want_scores=False needs want_mask=True)")
        self.scores = None if not want_scores else (scores if scores is not None else torch.empty(B, C, H, W, device=self.dev))
        self.mask = torch.empty(B, H, W, device=self.dev, dtype=torch.int64) if want_mask else None
        self.live = []               # (tensors kept alive, fill-done event) of the batches in flight
        self.fill_events = []        # optional (start, end) timing events of the fill kernel
        torch.cuda.current_stream(self.dev).synchronize()

    def submit(self, x, xs, pred, time_fill=False):
        g, R = self.g, self.R
        cur = torch.cuda.current_stream(self.dev)
        ready = torch.cuda.Event()
        ready.record(cur)                                    # inputs are produced on the caller's stream
        with torch.cuda.stream(self.plan_stream):
            self.plan_stream.wait_event(ready)
            if len(self.live) >= self.depth:                 # bound the run-ahead (and the memory held by plans)
                self.plan_stream.wait_event(self.live[0][1])
            grid = ops.saliency_to_grid(xs, self.g1x, self.g1y, g, g, R, R, "replication", (g, g))
            x_sampled = ops.grid_sample(x, grid)
            sampled = torch.cuda.Event()
            sampled.record(self.plan_stream)
            if self.interp == "nearest":
                plan = ops.build_nearest_plan(grid, (self.H, self.W), nchan=self.C)
            else:
                plan = ops.build_inverse_plan(grid, (self.H, self.W), nchan=self.C, triangulation=self.tri,
                                              dense_winner=False)
            table = ops.box4_table(pred)                     # A8 at the nodes: also off the fill stream, which then
            planned = torch.cuda.Event()                     # carries nothing but back-to-back fills
            planned.record(self.plan_stream)
        with torch.cuda.stream(self.fill_stream):
            self.fill_stream.wait_event(planned)             # (plan / table buffers stay referenced in self.live until
                                                             # `depth` batches later: no record_stream, whose deferred
                                                             # frees make the caching allocator grow and cudaMalloc)
            if time_fill:
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record(self.fill_stream)
            ops._fill(plan, table, self.C, True, self.scores, self.mask)
            if time_fill:
                t1.record(self.fill_stream)
                self.fill_events.append((t0, t1))
            done = torch.cuda.Event()
            done.record(self.fill_stream)
        # Contract: x_sampled is ordered on the caller's stream before it is returned (the encoder can consume it like any
        # torch result); the inputs x, xs, pred are read on the pipeline's streams AFTER submit returns, so references are
        # held until `depth` batches later -- the caching allocator cannot recycle them under the kernels even if the
        # caller drops them (they must not be overwritten in place before then).
        cur.wait_event(sampled)
        self.live.append(((plan, grid, table, x_sampled, x, xs, pred), done))
        if len(self.live) > self.depth:
            self.live.pop(0)
        return x_sampled, (self.scores if self.scores is not None else self.mask)

    def fence(self, stream=None):
        stream = stream or torch.cuda.current_stream(self.dev)
        stream.wait_stream(self.plan_stream)
        stream.wait_stream(self.fill_stream)

    def check(self):
        """Raise if the device Delaunay kernel reported a non-converged frame in any batch still referenced (host sync)."""
        for item, _ in self.live:
            ops.check_plan(item[0])

"""One training step of the hot path as a user-facing object: SQL decoder tail -> photometric losses ->
backward, on static device buffers, optionally captured in a CUDA graph (the step is ~100 small launches
around 10 large ones, so eager submission is CPU-launch-bound on a B200).

What a step covers (reference lines): networks/depth_decoder_QTR.py:47-70, trainer.py:386-439 (generate_images_pred)
and trainer.py:455-549 (compute_losses), then autograd back to the decoder features x, the queries, the 1x1
conv / bins-MLP weights, the pose parameters and the coarser-scale depth maps.
"""
import torch

from . import photometric as P
from . import sql as S


class HotPathConfig:
    def __init__(self, B=12, H=192, W=640, h=96, w=320, E=32, Q=64, D=64, S=2, scales=(0, 1, 2, 3),
                 min_depth=0.001, max_depth=80.0, disparity_smoothness=1e-3, stereo=False, automask=True):
        self.B, self.H, self.W, self.h, self.w = B, H, W, h, w
        self.E, self.Q, self.D, self.S = E, Q, D, S
        self.scales = tuple(scales)
        self.min_depth, self.max_depth = min_depth, max_depth
        self.disparity_smoothness = disparity_smoothness
        # --use_stereo (trainer.py:52-53, 408-421): the LAST source is the stereo frame "s" with the fixed transform
        # inputs["stereo_T"], and the PoseCNN translations are not rescaled by mean(1/depth)
        self.stereo = bool(stereo)
        self.automask = bool(automask)          # False: --disable_automasking (trainer.py:480,514,520)

    @property
    def pose_sources(self):
        """indices of the sources whose pose comes from the pose network (all but the stereo frame)"""
        return list(range(self.S - 1 if self.stereo else self.S))

    def scale_hw(self, s):
        """resolution of outputs[("disp", s)]: the decoder emits scale 0 at h x w; coarser scales are H/2^s."""
        return (self.h, self.w) if s == 0 else (self.H // 2 ** s, self.W // 2 ** s)

    def input_shapes(self):
        c = self
        shp = {"x": (c.B, c.E, c.h, c.w), "queries": (c.B, c.Q, c.E), "K": (c.B, 4, 4), "inv_K": (c.B, 4, 4),
               "target": (c.B, 3, c.H, c.W)}
        for i in range(c.S):
            shp["source%d" % i] = (c.B, 3, c.H, c.W)
        for i in c.pose_sources:
            shp["axisangle%d" % i] = (c.B, 1, 1, 3)
            shp["translation%d" % i] = (c.B, 1, 1, 3)
        if c.stereo:
            shp["stereo_T"] = (c.B, 4, 4)
        for s in c.scales:
            shp["noise%d" % s] = (c.B, c.S, c.H, c.W)
            if s > 0:
                shp["disp%d" % s] = (c.B, 1) + c.scale_hw(s)
                shp["target%d" % s] = (c.B, 3, c.H // 2 ** s, c.W // 2 ** s)
        return shp


class HotPath(torch.nn.Module):
    """Parameters: the decoder's 1x1 conv (convert_to_prob.0) and bins_regressor.  Inputs live in static
    device buffers `self.inp[name]` (see HotPathConfig.input_shapes); `load()` copies a host batch into them."""

    def __init__(self, cfg, device="cuda", use_graph=True, num_slots=1, grad_exchange=None, exchange_sm_reserve=32,
                 prepare_next=True, prepare_fork="auto"):
        super().__init__()
        nn = torch.nn
        self.cfg = cfg
        c = cfg
        self.convert_to_prob = nn.Sequential(nn.Conv2d(c.Q, c.D, kernel_size=1), nn.Softmax(dim=1))
        self.bins_regressor = nn.Sequential(nn.Linear(c.E * c.Q, 16 * c.Q), nn.LeakyReLU(),
                                            nn.Linear(16 * c.Q, 256), nn.LeakyReLU(), nn.Linear(256, c.D))
        self.to(device)
        self.device = torch.device(device)
        self.grad_inputs = ["x", "queries"] + ["disp%d" % s for s in c.scales if s > 0] + \
                           ["axisangle%d" % i for i in c.pose_sources] + ["translation%d" % i for i in c.pose_sources]
        # `num_slots` independent device input sets: with 2, the host->device copy of batch i+1 (on a copy stream)
        # overlaps the step on batch i, as a pinned-memory DataLoader does for the reference (trainer.py:164-171)
        # Every input of a set is a view into ONE flat device buffer (frames in a second one), so a host batch that
        # is packed the same way crosses PCIe as two large copies instead of a dozen small ones (load_flat()).
        shapes = cfg.input_shapes()
        self.frame_keys = [k for k in shapes if k.startswith("target") or k.startswith("source")]
        # the tie-break noise planes go last: a caller that draws them on the device ships only the leading part
        self.other_keys = [k for k in shapes if k not in self.frame_keys and not k.startswith("noise")] + \
                          [k for k in shapes if k.startswith("noise")]

        def layout(keys):
            off, table = 0, {}
            for k in keys:
                n = 1
                for d in shapes[k]:
                    n *= d
                table[k] = (off, n)
                off += (n + 63) // 64 * 64          # 256-byte aligned views
            return table, off
        self.frame_layout, self.frame_numel = layout(self.frame_keys)
        self.other_layout, self.other_numel = layout(self.other_keys)
        self.slots, self.flat = [], []
        for _ in range(num_slots):
            fl = {"frames": torch.zeros(self.frame_numel, device=device, dtype=torch.float32),
                  "frames_u8": torch.zeros(self.frame_numel, device=device, dtype=torch.uint8),
                  "other": torch.zeros(self.other_numel, device=device, dtype=torch.float32)}
            inp = {}
            for k in self.frame_keys:
                o, n = self.frame_layout[k]
                inp[k] = fl["frames"].narrow(0, o, n).view(shapes[k])
            for k in self.other_keys:
                o, n = self.other_layout[k]
                inp[k] = fl["other"].narrow(0, o, n).view(shapes[k])
            for k in self.grad_inputs:
                inp[k].requires_grad_(True)
            self.slots.append(inp)
            self.flat.append(fl)
        self.inp = self.slots[0]
        # ONE flat gradient bucket per input set: the backward kernels write every parameter gradient of the path
        # (1x1 conv, bins MLP) straight into 256-byte aligned views of it, `p.grad` are those views, and the
        # data-parallel exchange all-reduces the flat tensor in place -- no pack / unpack copies, no per-parameter
        # zero-fill or accumulate kernels (the reference's nn.DataParallel reduce-adds replica gradients, trainer.py:74,93)
        self.param_list = list(self.parameters())
        offs, total = [], 0
        for p_ in self.param_list:
            offs.append(total)
            total += (p_.numel() + 63) // 64 * 64
        self.grad_flat = [torch.zeros(total, device=device, dtype=torch.float32) for _ in range(num_slots)]
        self.grad_views = [[fl.narrow(0, o, p_.numel()).view_as(p_) for o, p_ in zip(offs, self.param_list)]
                           for fl in self.grad_flat]
        self._one = torch.ones((), device=device, dtype=torch.float32)
        self.use_graph = use_graph
        self.graphs = [None] * num_slots
        self.losses = [None] * num_slots
        self.loss = None
        self.pred = None
        self.side_stream = None
        # Frame-only work of a step -- the identity reprojection losses (trainer.py:480-493) and the pixel-interleaved
        # source copies: 87 us of full-GPU kernels at config 2 that depend on nothing but the input frames.  Run at the
        # start of the step they compete with the decoder tail for SMs and barely overlap (measured: removing them takes
        # the step from 1.039 to 0.955 ms).  With prepare_next, every step instead computes them for the NEXT input set
        # (the one after `slot`, round-robin; the same set when there is one) on the side stream during its own backward
        # tail (forked right after the regression-path kernel: partial sums, bins-head backward, glue kernels follow, ~70 us
        # of grids that leave most SMs idle) into
        # persistent per-set buffers, and consumes what the previous step prepared.  Every step still does this work
        # exactly once, inside its own graph; a set whose frames changed after it was prepared (load) is re-prepared
        # eagerly before its step.
        self.prepare_next = prepare_next
        # Where the preparation is forked.  The step is captured from a high-priority stream and the preparation runs at
        # priority 0, but priorities only arbitrate FREE resources: a one-CTA-per-SM tcgen05 kernel (227 KB of shared
        # memory) cannot displace a stream of small identity-loss CTAs that keep refilling the SM, so the preparation must
        # not start while such kernels are still to come soon.  Measured on one B200 (ms/step, configs 2 / 3 / 4; frame-only
        # work at the start of its own step as in round 1: 1.039 / 1.095 / 3.033):
        #   "after_pred_fwd"  (the forward's last tcgen05 kernel is enqueued; the work then fills the wave tails of the
        #                     photometric kernels, the reduction tails and the single-block pose backward)  1.001 / 1.024 / 2.928
        #   "after_bwd_pred"  (~70 us of small grids follow)                                                 1.002 / 1.079 / 3.025
        #   "start"                                                                                          1.010 / 1.051 / 2.982
        #   end of the forward 1.043 / 1.051 / 3.052;  start of the decoder-tail backward 1.031 / 1.095 / 3.033
        # With an in-step gradient exchange the window after the regression-path backward belongs to the collective
        # (2 x B200, config 2: "after_bwd_pred" 1.051 ms with 49 us of exposed communication, "start" 1.039 ms with 15 us).
        # "auto" = "after_pred_fwd"; preparing the SAME set (one input set) always waits for "after_bwd_pred".
        assert prepare_fork in ("auto", "start", "after_pred_fwd", "after_bwd_pred")
        self.prepare_fork = prepare_fork
        self._prep = [None] * num_slots            # (identity [B,S,H,W] or None, [S x packed [B,H,W,4]])
        self._prep_valid = [False] * num_slots
        self._prep_forked = False
        self._capture_stream = None
        self._graph_prepares = [False] * num_slots   # whether the captured graph of a set prepares the next one
        # grad_exchange(flat gradient bucket): in-place data-parallel exchange (e.g. an NCCL all-reduce, average) of the
        # parameter gradients.  It is issued INSIDE the step, on a side stream, as soon as the last parameter gradient
        # exists, so that it overlaps the summary-path backward kernel; the step (eager or captured graph, NCCL
        # collectives are capturable) ends with the main stream waiting for it.
        self.grad_exchange = grad_exchange
        # SMs the summary-path backward leaves to the communication kernel while the exchange runs beside it (measured on
        # 8 x B200, NCCL 2.28: exposed communication 88 us with 0, 35 us with 32, 53 us with 64; tools/nccl_sweep.sh)
        self.exchange_sm_reserve = exchange_sm_reserve
        self.comm_stream = None

    # ------------------------------------------------------------------ data
    def load(self, host_batch, non_blocking=True, slot=0):
        """host_batch: {name: CPU tensor (pinned for async copies)} -> device input set `slot` on the current
        stream; returns the bytes copied.  Frames may be shipped as uint8 (what the dataset decodes to): they cross
        PCIe at one byte per channel and are scaled to [0,1] floats on the device -- the reference does the same
        conversion (transforms.ToTensor) on the host and ships four bytes per channel."""
        n = 0
        inp = self.slots[slot]
        self._prep_valid[slot] = False
        with torch.no_grad():
            for k, v in host_batch.items():
                inp[k].copy_(v, non_blocking=non_blocking)
                if v.dtype == torch.uint8:
                    inp[k].mul_(1.0 / 255.0)
                n += v.numel() * v.element_size()
        return n

    def pack_host(self, host_batch, pin=True):
        """Pack a host batch dict into the flat layout of the device input sets: (frames, other) where `frames` is
        uint8 when the batch carries uint8 frames (else float32) and `other` is float32.  A data loader would
        collate straight into these buffers."""
        u8 = all(host_batch[k].dtype == torch.uint8 for k in self.frame_keys)
        frames = torch.zeros(self.frame_numel, dtype=torch.uint8 if u8 else torch.float32)
        other = torch.zeros(self.other_numel, dtype=torch.float32)
        for k in self.frame_keys:
            o, n = self.frame_layout[k]
            frames[o:o + n] = host_batch[k].reshape(-1)
        for k in self.other_keys:
            o, n = self.other_layout[k]
            other[o:o + n] = host_batch[k].reshape(-1).float()
        if pin:
            frames, other = frames.pin_memory(), other.pin_memory()
        return frames, other

    def other_numel_without(self, prefixes=("noise",)):
        """Length of the leading part of the flat `other` buffer that excludes the trailing keys with these prefixes."""
        first = [self.other_layout[k][0] for k in self.other_keys if k.startswith(tuple(prefixes))]
        return min(first) if first else self.other_numel

    def noise_region(self, slot=0):
        """The contiguous tail of the flat buffer holding every scale's tie-break noise (one normal_() fills it)."""
        return self.flat[slot]["other"][self.other_numel_without(("noise",)):]

    def load_flat(self, frames, other, non_blocking=True, slot=0):
        """Two host->device copies for a whole batch (pack_host layout) plus, for uint8 frames, ONE scaling kernel;
        returns the bytes copied."""
        fl = self.flat[slot]
        self._prep_valid[slot] = False
        with torch.no_grad():
            fl["other"][:other.numel()].copy_(other, non_blocking=non_blocking)
            if frames.dtype == torch.uint8:
                fl["frames_u8"].copy_(frames, non_blocking=non_blocking)
                torch.mul(fl["frames_u8"], 1.0 / 255.0, out=fl["frames"])
            else:
                fl["frames"].copy_(frames, non_blocking=non_blocking)
        return frames.numel() * frames.element_size() + other.numel() * other.element_size()

    # ------------------------------------------------------------------ one eager step
    def _centers_fn(self, slot):
        c = self.cfg
        bufs = self.grad_views[slot][2:]           # [dW1, db1, dW2, db2, dW3, db3]

        def centers(summary):
            return S.bins_head(summary.reshape(c.B, c.Q * c.E), self.bins_regressor, c.min_depth, c.max_depth,
                               grad_buffers=bufs)
        return centers

    def forward_loss(self, slot=0):
        c, I = self.cfg, self.slots[slot]
        conv = self.convert_to_prob[0]
        sources = [I["source%d" % i] for i in range(c.S)]
        # what depends on the input frames only (identity losses, pixel-interleaved source copies) runs on a side
        # stream while the decoder tail -- a few 4-warp CTAs per SM -- occupies the main one
        main = torch.cuda.current_stream()
        if self.side_stream is None:
            self.side_stream = torch.cuda.Stream()
        side = self.side_stream
        if self.prepare_next:
            # prepared by the previous step (or eagerly by step() / step_eager() when the set's frames are new)
            identity, packed = self._prep[slot]
            side = None
        else:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                identity = P.identity_losses(I["target"], sources) if c.automask else None
                packed = [P.pack_rgba(src) for src in sources]
        gv = self.grad_views[slot]
        hook = (lambda grads: self._start_grad_exchange(slot)) if self.grad_exchange is not None else None
        S.exchange_sm_reserve = self.exchange_sm_reserve if hook is not None else 0
        nxt = (slot + 1) % len(self.slots)
        stage = None
        if self.prepare_next:
            where = self.prepare_fork
            if where == "auto":
                where = "after_pred_fwd"
            if nxt == slot or not torch.is_grad_enabled():
                where = "after_bwd_pred"       # the SAME set: its packed sources are read until the photometric backward is done
            if where == "start":
                self._fork_prepare(nxt)        # another input set: its buffers are independent of this step
            else:
                stage = lambda name: self._fork_prepare(nxt) if name == where else None  # noqa: E731
        pred = S.sql_tail(I["x"], I["queries"], conv.weight.view(c.D, c.Q), conv.bias, self._centers_fn(slot), (),
                          on_param_grads=hook, head_grad_out=(gv[0].view(c.D, c.Q), gv[1]), on_stage=stage)
        if side is not None:
            main.wait_stream(side)
        disps = {s: (pred if s == 0 else I["disp%d" % s]) for s in c.scales}
        target_pyr = {s: (I["target"] if s == 0 else I["target%d" % s]) for s in c.scales}
        poses = [{"axisangle": I["axisangle%d" % i], "translation": I["translation%d" % i], "invert": i == 0}
                 for i in c.pose_sources]
        if c.stereo:
            poses.append({"T": I["stereo_T"]})
        noises = {s: I["noise%d" % s] for s in c.scales}
        out = P.photometric_losses(disps, target_pyr, sources, I["K"], I["inv_K"], poses, noises, height=c.H,
                                   width=c.W, scales=c.scales, disparity_smoothness=c.disparity_smoothness,
                                   rescale_translation=not c.stereo, disable_automasking=not c.automask,
                                   identity=identity, packed_sources=packed)
        self.argmins = {s: out[("argmin", s)] for s in c.scales}       # [B,H,W] u8 per scale (what Trainer.log reads)
        return out["loss"], pred

    def _start_grad_exchange(self, slot):
        """fork: the exchange of the slot's flat gradient bucket runs on the communication stream after everything
        enqueued so far; returns the join"""
        main = torch.cuda.current_stream()
        if self.comm_stream is None:
            # high priority, like the step itself: the collective's CTAs spin on their peers and must not queue behind
            # the low-priority preparation of the next input set
            self.comm_stream = torch.cuda.Stream(priority=-1)
        comm = self.comm_stream
        comm.wait_stream(main)
        with torch.cuda.stream(comm):
            self.grad_exchange(self.grad_flat[slot])
        return lambda: torch.cuda.current_stream().wait_stream(comm)

    # ------------------------------------------------------------------ frame-only work (see prepare_next)
    def _run_prepare(self, slot):
        """identity losses + packed sources of input set `slot` into its persistent buffers, on the current stream"""
        c, I = self.cfg, self.slots[slot]
        sources = [I["source%d" % i] for i in range(c.S)]
        if self._prep[slot] is None:
            ident = torch.empty(c.B, c.S, c.H, c.W, device=self.device, dtype=torch.float32) if c.automask else None
            packed = [torch.empty(c.B, c.H, c.W, 4, device=self.device, dtype=torch.float32) for _ in sources]
            self._prep[slot] = (ident, packed)
        ident, packed = self._prep[slot]
        with torch.no_grad():
            if ident is not None:
                P.identity_losses(I["target"], sources, out=ident)
            for src, out in zip(sources, packed):
                P.pack_rgba(src, out=out)

    def _fork_prepare(self, slot):
        """backward hook: the photometric backward (the last reader of the packed sources) is enqueued; fork the
        preparation of the next input set onto the side stream, under the decoder tail's backward"""
        main = torch.cuda.current_stream()
        if self.side_stream is None:
            self.side_stream = torch.cuda.Stream()
        self.side_stream.wait_stream(main)
        with torch.cuda.stream(self.side_stream):
            self._run_prepare(slot)
        self._prep_forked = True
        return None

    def _join_prepare(self):
        """join the forked preparation (if the backward hook fired); returns whether there was one"""
        forked, self._prep_forked = self._prep_forked, False
        if forked:
            torch.cuda.current_stream().wait_stream(self.side_stream)
        return forked

    def _ensure_prepared(self, slot):
        """eager preparation of a set whose frames were (re)loaded after the last step prepared it"""
        if self.prepare_next and not self._prep_valid[slot]:
            self._run_prepare(slot)
            self._prep_valid[slot] = True

    def _zero_grads(self, slot=0):
        for k in self.grad_inputs:
            self.slots[slot][k].grad = None

    def _point_grads(self, slot):
        for p, gv in zip(self.param_list, self.grad_views[slot]):
            p.grad = gv

    def step_eager(self, slot=0):
        self._zero_grads(slot)
        self._ensure_prepared(slot)
        loss, pred = self.forward_loss(slot)
        torch.autograd.backward(loss, grad_tensors=self._one)     # (a static one: no fill kernel per step)
        if self._join_prepare():
            self._prep_valid[(slot + 1) % len(self.slots)] = True
        self._point_grads(slot)
        self.loss, self.pred = loss.detach(), pred.detach()
        self.losses[slot] = self.loss
        return self.loss

    # ------------------------------------------------------------------ graph
    def capture(self, warmup=3, slot=0):
        """Capture forward + backward on input set `slot` into one CUDA graph (parameter gradients land in the slot's
        flat bucket; input gradients are the tensors autograd creates during capture: static addresses in the pool)."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.step_eager(slot)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for k in self.grad_inputs:
            self.slots[slot][k].grad = None
        if self.prepare_next:                       # the persistent buffers must exist before the capture
            for sl in {slot, (slot + 1) % len(self.slots)}:
                if self._prep[sl] is None:
                    self._run_prepare(sl)
                    self._prep_valid[sl] = True
            torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        # the step is captured from a HIGH-priority stream: the kernel nodes keep that priority, so whenever SM resources
        # free up the step's own chain wins them over the side-stream work (next set's preparation, priority 0)
        if self._capture_stream is None:
            self._capture_stream = torch.cuda.Stream(priority=-1)
        with torch.cuda.graph(g, stream=self._capture_stream):
            loss, pred = self.forward_loss(slot)
            torch.autograd.backward(loss, grad_tensors=self._one)
            self._graph_prepares[slot] = self._join_prepare()
            self.losses[slot] = loss.detach()
            self.pred = pred.detach()
        self.graphs[slot] = g
        return g

    def drop_graphs(self):
        """Release the captured graphs (before tearing down a process group whose collectives they hold)."""
        self.graphs = [None] * len(self.graphs)

    def step(self, slot=0):
        """Run one step on whatever is in input set `slot`; returns the (device) loss scalar."""
        if self.use_graph:
            if self.graphs[slot] is None:
                self.capture(slot=slot)
            self._ensure_prepared(slot)
            self.graphs[slot].replay()
            if self._graph_prepares[slot]:
                self._prep_valid[(slot + 1) % len(self.slots)] = True
            self._point_grads(slot)
            self.loss = self.losses[slot]
            return self.loss
        return self.step_eager(slot)

    def param_grads(self):
        return [p.grad for p in self.param_list if p.grad is not None]

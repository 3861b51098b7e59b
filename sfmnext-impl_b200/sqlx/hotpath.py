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
                 min_depth=0.001, max_depth=80.0, disparity_smoothness=1e-3):
        self.B, self.H, self.W, self.h, self.w = B, H, W, h, w
        self.E, self.Q, self.D, self.S = E, Q, D, S
        self.scales = tuple(scales)
        self.min_depth, self.max_depth = min_depth, max_depth
        self.disparity_smoothness = disparity_smoothness

    def scale_hw(self, s):
        """resolution of outputs[("disp", s)]: the decoder emits scale 0 at h x w; coarser scales are H/2^s."""
        return (self.h, self.w) if s == 0 else (self.H // 2 ** s, self.W // 2 ** s)

    def input_shapes(self):
        c = self
        shp = {"x": (c.B, c.E, c.h, c.w), "queries": (c.B, c.Q, c.E), "K": (c.B, 4, 4), "inv_K": (c.B, 4, 4),
               "target": (c.B, 3, c.H, c.W)}
        for i in range(c.S):
            shp["source%d" % i] = (c.B, 3, c.H, c.W)
            shp["axisangle%d" % i] = (c.B, 1, 1, 3)
            shp["translation%d" % i] = (c.B, 1, 1, 3)
        for s in c.scales:
            shp["noise%d" % s] = (c.B, c.S, c.H, c.W)
            if s > 0:
                shp["disp%d" % s] = (c.B, 1) + c.scale_hw(s)
                shp["target%d" % s] = (c.B, 3, c.H // 2 ** s, c.W // 2 ** s)
        return shp


class HotPath(torch.nn.Module):
    """Parameters: the decoder's 1x1 conv (convert_to_prob.0) and bins_regressor.  Inputs live in static
    device buffers `self.inp[name]` (see HotPathConfig.input_shapes); `load()` copies a host batch into them."""

    def __init__(self, cfg, device="cuda", use_graph=True, num_slots=1, grad_exchange=None):
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
                           ["axisangle%d" % i for i in range(c.S)] + ["translation%d" % i for i in range(c.S)]
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
        self.use_graph = use_graph
        self.graphs = [None] * num_slots
        self.losses = [None] * num_slots
        self.loss = None
        self.pred = None
        # parameter gradients of each captured graph live in that graph's memory pool (autograd ASSIGNS them: no
        # zero-fill + accumulate kernels per parameter); step(slot) points p.grad at the replayed graph's tensors
        self.slot_grads = [None] * num_slots
        self.side_stream = None
        # grad_exchange(list of gradient tensors): in-place data-parallel exchange (e.g. GradBucket.allreduce_) of the
        # parameter gradients.  It is issued INSIDE the step, on a side stream, as soon as the last parameter gradient
        # exists, so that it overlaps the summary-path backward kernel; the step (eager or captured graph, NCCL
        # collectives are capturable) ends with the main stream waiting for it.
        self.grad_exchange = grad_exchange
        self.comm_stream = None

    # ------------------------------------------------------------------ data
    def load(self, host_batch, non_blocking=True, slot=0):
        """host_batch: {name: CPU tensor (pinned for async copies)} -> device input set `slot` on the current
        stream; returns the bytes copied.  Frames may be shipped as uint8 (what the dataset decodes to): they cross
        PCIe at one byte per channel and are scaled to [0,1] floats on the device -- the reference does the same
        conversion (transforms.ToTensor) on the host and ships four bytes per channel."""
        n = 0
        inp = self.slots[slot]
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
        with torch.no_grad():
            fl["other"][:other.numel()].copy_(other, non_blocking=non_blocking)
            if frames.dtype == torch.uint8:
                fl["frames_u8"].copy_(frames, non_blocking=non_blocking)
                torch.mul(fl["frames_u8"], 1.0 / 255.0, out=fl["frames"])
            else:
                fl["frames"].copy_(frames, non_blocking=non_blocking)
        return frames.numel() * frames.element_size() + other.numel() * other.element_size()

    # ------------------------------------------------------------------ one eager step
    def _centers(self, summary):
        c = self.cfg
        return S.bins_head(summary.reshape(c.B, c.Q * c.E), self.bins_regressor, c.min_depth, c.max_depth)

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
        side.wait_stream(main)
        with torch.cuda.stream(side):
            identity = P.identity_losses(I["target"], sources)
            packed = [P.pack_rgba(src) for src in sources]
        pred = S.sql_tail(I["x"], I["queries"], conv.weight.view(c.D, c.Q), conv.bias, self._centers,
                          tuple(self.bins_regressor.parameters()),
                          on_param_grads=self._start_grad_exchange if self.grad_exchange is not None else None)
        main.wait_stream(side)
        disps = {s: (pred if s == 0 else I["disp%d" % s]) for s in c.scales}
        target_pyr = {s: (I["target"] if s == 0 else I["target%d" % s]) for s in c.scales}
        poses = [{"axisangle": I["axisangle%d" % i], "translation": I["translation%d" % i], "invert": i == 0}
                 for i in range(c.S)]
        noises = {s: I["noise%d" % s] for s in c.scales}
        out = P.photometric_losses(disps, target_pyr, sources, I["K"], I["inv_K"], poses, noises, height=c.H,
                                   width=c.W, scales=c.scales, disparity_smoothness=c.disparity_smoothness,
                                   identity=identity, packed_sources=packed)
        return out["loss"], pred

    def _start_grad_exchange(self, grads):
        """fork: the exchange runs on the communication stream after everything enqueued so far; returns the join"""
        main = torch.cuda.current_stream()
        if self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream()
        comm = self.comm_stream
        comm.wait_stream(main)
        with torch.cuda.stream(comm):
            self.grad_exchange(grads)
        return lambda: torch.cuda.current_stream().wait_stream(comm)

    def _zero_grads(self, slot=0):
        for p in self.parameters():
            p.grad = None
        for k in self.grad_inputs:
            self.slots[slot][k].grad = None

    def step_eager(self, slot=0):
        self._zero_grads(slot)
        loss, pred = self.forward_loss(slot)
        loss.backward()
        self.loss, self.pred = loss.detach(), pred.detach()
        self.losses[slot] = self.loss
        return self.loss

    # ------------------------------------------------------------------ graph
    def capture(self, warmup=3, slot=0):
        """Capture forward + backward on input set `slot` into one CUDA graph (parameter and input gradients are the
        tensors autograd creates during capture: static addresses in the graph's pool)."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.step_eager(slot)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for k in self.grad_inputs:
            self.slots[slot][k].grad = None
        g = torch.cuda.CUDAGraph()
        for p in self.parameters():
            p.grad = None
        with torch.cuda.graph(g):
            loss, pred = self.forward_loss(slot)
            loss.backward()
            self.losses[slot] = loss.detach()
            self.pred = pred.detach()
        self.slot_grads[slot] = [p.grad for p in self.parameters()]
        self.graphs[slot] = g
        return g

    def step(self, slot=0):
        """Run one step on whatever is in input set `slot`; returns the (device) loss scalar."""
        if self.use_graph:
            if self.graphs[slot] is None:
                self.capture(slot=slot)
            self.graphs[slot].replay()
            for p, gr in zip(self.parameters(), self.slot_grads[slot]):
                p.grad = gr
            self.loss = self.losses[slot]
            return self.loss
        return self.step_eager(slot)

    def param_grads(self):
        return [p.grad for p in self.parameters() if p.grad is not None]

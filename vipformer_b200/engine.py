"""Data-parallel pre-training step: the hot loop of pretrain.py:173-211, one process per GPU.

    for ((pc_t1, pc_t2), imgs) in loader:  forward both branches -> NT-Xent (intra + cross modal) -> backward
                                           -> gradient all-reduce (DDP, pretrain.py:104-105) -> AdamW (:121-124,:210)

B200-first structure:
  * both models live in ONE parameter arena: a single flat fp32 gradient buffer is zeroed by one memset, all-reduced
    by ONE NCCL call over NVLink/NVSwitch, and consumed by ONE fused AdamW kernel that also refreshes the bf16 shadows;
  * the whole step (FPS start draw, ~480 kernel launches on three streams, collectives, optimizer) is captured in a CUDA graph and
    replayed; per-step dropout seeds / step count / learning rate live in device memory so the graph stays valid;
  * with world_size > 1 NT-Xent negatives span the global batch (all-gather of normalised embeddings + of the per-row
    log-sum-exp); BatchNorm statistics stay rank-local exactly like the reference (no SyncBN);
  * GradScaler (pretrain.py:154,209-211) is dropped: bf16 operands with fp32 accumulation need no loss scaling.
"""
import os
import weakref

import torch
import torch.nn as nn

from . import ops, params, runtime
from .loss import pretrain_loss

FPS_START_OP = 0xF9500000   # stream id of the per-step FPS start draw (ops.draw_indices); dropout sites use small ids

F32 = torch.float32


def allreduce_gradients(root, op=None):
    """DDP's job for a hand-written loop around the mirror modules (pretrain.py:104-105): average the flat gradient
    buffer of `root`'s parameter arena over the ranks with ONE NCCL all-reduce.  torch's DistributedDataParallel cannot
    wrap the mirror: the block Functions write parameter gradients straight into the arena (p.grad views), so DDP's
    per-parameter autograd hooks never fire.  Call between loss.backward() and optimizer.step()."""
    import torch.distributed as dist

    arena = params.arena_of(root)
    if arena.flat_g is None:
        raise RuntimeError("no gradients yet: call after backward()")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(arena.flat_g, op=dist.ReduceOp.AVG if op is None else op)
    return arena.flat_g


class PretrainEngine:
    def __init__(self, pc_model, img_model, *, batch_pairs, num_points, img_size=144, lr=1e-3, betas=(0.9, 0.999),
                 eps=1e-8, weight_decay=1e-2, temperature=0.1, cmid_weight=1.0, gather_distributed=None,
                 use_cuda_graph=True, seed=0, device=None, images_nchw=True, overlap_branches=True):
        import torch.distributed as dist

        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.dist = dist if (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) else None
        self.world = self.dist.get_world_size() if self.dist else 1
        self.rank = self.dist.get_rank() if self.dist else 0
        self.gather = (self.world > 1) if gather_distributed is None else bool(gather_distributed)
        self.pc_model, self.img_model = pc_model.to(self.device).train(), img_model.to(self.device).train()
        self.root = nn.ModuleList([self.pc_model, self.img_model])
        for m in self.root.modules():
            object.__setattr__(m, "_vpf_root", self.root) if m is not self.root else None
        if self.dist:   # DDP wrap-time parameter broadcast (pretrain.py:104-105)
            for p in self.root.parameters():
                self.dist.broadcast(p.data, src=0)
            for b in self.root.buffers():
                self.dist.broadcast(b.data, src=0)
        # dropout site ids in module order, so a run depends on (weights, seed) only and not on how many layers the
        # process constructed before
        for i, m in enumerate(m for m in self.root.modules() if hasattr(m, "_op_base")):
            m._op_base = 8 * (i + 1)
        self.arena = params.prepare(self.root, self.device)
        self.arena.ensure_grads()
        self.arena.refresh_shadows(force=True)
        self.arena.managed = True
        n = self.arena.flat_p.numel()
        n_pc_params = len(list(self.pc_model.parameters()))
        self.n_pc = self.arena._offsets[n_pc_params] if n_pc_params < len(self.arena._offsets) else n   # first image-model element
        self.m = ops.zeros_(torch.empty(n, dtype=F32, device=self.device))
        self.v = ops.zeros_(torch.empty(n, dtype=F32, device=self.device))
        self.lr = torch.tensor([lr], dtype=F32, device=self.device)
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay
        self.temperature, self.cmid_weight = temperature, cmid_weight
        self.state = runtime.StepState.get(self.device)   # one engine per process/GPU, like the reference's mp.spawn
        self.state[0] = 0                                  # optimizer step count (AdamW bias correction)
        runtime.manual_seed(seed * 1000003 + self.rank, self.device)
        b = batch_pairs
        self.b, self.N = b, num_points
        self.images_nchw = images_nchw
        # static step inputs (graph-stable addresses); [t1; t2] is the concatenation of pretrain.py:183
        self.pc_in = torch.zeros((2 * b, num_points, 3), dtype=F32, device=self.device)
        ishape = (b, 3, img_size, img_size) if images_nchw else (b, img_size, img_size, 3)
        self.img_in = torch.zeros(ishape, dtype=F32, device=self.device)
        self.start = torch.zeros(2 * b, dtype=torch.long, device=self.device)
        self.losses = torch.zeros(3, dtype=F32, device=self.device)
        self.losses_host = torch.zeros(3, dtype=F32).pin_memory()
        self.use_graph = use_cuda_graph
        self.graph = None
        self.side = torch.cuda.Stream(device=self.device) if overlap_branches else None
        self.overlap_allreduce = os.environ.get("VPF_AR_OVERLAP", "1") != "0"
        self._img_after_g2e = self.side is not None and os.environ.get("VPF_IMG_AFTER_G2E", "1") != "0"
        if self._img_after_g2e:
            self._g2e_done = None      # armed only inside _step_body: the hook is inert when the model is called on its own
            me = weakref.ref(self)     # the module must not keep the engine (and its CUDA graph with captured NCCL kernels) alive

            def _mark(module, inputs, output):
                eng = me()
                if eng is not None and eng._g2e_done is not None:
                    eng._g2e_done.record()

            self._g2e_hook = self.pc_model.group2emb.register_forward_hook(_mark)
        self._copy_stream, self._staged = None, False
        self._loss_ring, self._loss_pending = None, None
        self.steps_done = 0

    def close(self):
        """Drop the captured graph (it holds NCCL kernels: do this BEFORE destroy_process_group) and the scheduling hook."""
        self.graph = None
        h = self.__dict__.pop("_g2e_hook", None)
        if h is not None:
            h.remove()

    # ------------------------------------------------------------------------------------------------ one step
    def set_lr(self, lr):
        self.lr.fill_(float(lr))

    def _step_body(self):
        ops.step_advance(self.state)                      # step += 1, fresh dropout seed
        self.arena.zero_grads()                           # optimizer.zero_grad (pretrain.py:174)
        ops.draw_indices(self.state, FPS_START_OP, self.N, self.start)   # utils.py:71, drawn on the device from the step seed
        self.pc_model.fps_start_idx = self.start
        imgs = self.img_in.permute(0, 2, 3, 1) if self.images_nchw else self.img_in   # pretrain.py:179
        if self.side is not None:
            # the two encoders are independent until the loss: the image branch runs on a second stream so that its
            # kernels fill the partial last waves / launch gaps of the point-cloud branch (autograd replays each
            # node's backward on the stream its forward ran on, so the backward overlaps the same way)
            cur = torch.cuda.current_stream()
            if not self._img_after_g2e:
                self.side.wait_stream(cur)
                with torch.cuda.stream(self.side):
                    img_feats, _ = self.img_model(imgs)       # pretrain.py:199
                pc_feats, _ = self.pc_model(self.pc_in)       # pretrain.py:186
            else:
                # The image forward starts when the point-cloud branch leaves Group2Emb (event recorded by a forward hook):
                # the tokenizer + Group2Emb kernels are few and large and gain nothing from a neighbour, the ~170
                # 20-70 us kernels of the point-cloud encoder behind them do (their launch gaps and partial last waves get
                # filled).  Measured -0.1 ... -0.3 ms per step against starting both branches together.
                self._g2e_done = torch.cuda.Event()
                pc_feats, _ = self.pc_model(self.pc_in)
                with torch.cuda.stream(self.side):
                    self.side.wait_event(self._g2e_done)
                    img_feats, _ = self.img_model(imgs)
                self._g2e_done = None
            cur.wait_stream(self.side)
        else:
            pc_feats, _ = self.pc_model(self.pc_in)
            img_feats, _ = self.img_model(imgs)
        losses = pretrain_loss(pc_feats, img_feats, self.temperature, self.cmid_weight, self.gather)
        if self.dist and self.side is not None and self.overlap_allreduce:
            # DDP-style overlap without per-parameter hooks: backward in three pieces.  (1) the loss node alone; (2) the image
            # branch, launched from its own stream, followed at once by the all-reduce of ITS slice of the flat gradient
            # buffer (parameters are laid out [point-cloud model | image model]); (3) the point-cloud branch, whose longer
            # backward (Group2Emb, 2048-key cross-attention) hides that transfer.  Only the point-cloud slice is reduced
            # after the backward.  All collectives stay on the launching thread, so the step still captures as one graph.
            cur = torch.cuda.current_stream()
            d_pc, d_img = torch.autograd.grad(losses[0], [pc_feats, img_feats])
            self.side.wait_stream(cur)
            with torch.cuda.stream(self.side):
                torch.autograd.backward([img_feats], [d_img])
                work = self.dist.all_reduce(self.arena.flat_g[self.n_pc:], op=self.dist.ReduceOp.AVG, async_op=True)
            torch.autograd.backward([pc_feats], [d_pc])
            cur.wait_stream(self.side)
            self.dist.all_reduce(self.arena.flat_g[:self.n_pc], op=self.dist.ReduceOp.AVG)
            work.wait()
        else:
            losses[0].backward()                              # pretrain.py:209
            if self.side is not None:
                torch.cuda.current_stream().wait_stream(self.side)
            if self.dist:                                     # DDP gradient all-reduce (mean), one flat buffer
                self.dist.all_reduce(self.arena.flat_g, op=self.dist.ReduceOp.AVG)
        ops.adamw(self.arena.flat_p, self.arena.flat_g, self.m, self.v, self.arena.flat_bf, self.lr, self.state[:1],
                  self.betas[0], self.betas[1], self.eps, self.weight_decay)   # pretrain.py:210
        ops.add_scale(losses.detach(), None, 1.0, out=self.losses)

    def _snapshot(self):
        """Everything a step mutates besides the activations: parameters, Adam moments, step/seed state (which also drives
        the FPS start draw), BatchNorm buffers."""
        return ([t.clone() for t in (self.arena.flat_p, self.arena.flat_bf, self.m, self.v, self.state)],
                [b.clone() for b in self.root.buffers()])

    def _restore(self, snap):
        tensors, bufs = snap
        for dst, src in zip((self.arena.flat_p, self.arena.flat_bf, self.m, self.v, self.state), tensors):
            dst.copy_(src)
        for dst, src in zip(self.root.buffers(), bufs):
            dst.copy_(src)

    def _capture(self):
        # Warm-up outside capture (allocator pools, lazy module init, NCCL) must not train: the reference takes exactly
        # one optimizer step per batch (pretrain.py:209-211).  Snapshot, run the body twice, put everything back.
        snap = self._snapshot()
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self._step_body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._restore(snap)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._step_body()

    def step(self):
        """Run one optimisation step on whatever is in self.pc_in / self.img_in; returns the device tensor
        (total, loss_imid, loss_cmid) without synchronising."""
        if self.use_graph:
            if self.graph is None:
                self._capture()
            self.graph.replay()
        else:
            self._step_body()
        self.steps_done += 1
        return self.losses

    # ------------------------------------------------------------------------------- pipelined host input (e2e)
    def prefetch_host(self, pc_t1, pc_t2, imgs):
        """Start the H2D copy of the NEXT step's inputs (pinned host tensors, same shapes as step_host) on a copy
        stream into a staging buffer; it overlaps whatever the compute stream is running.  What a DataLoader with
        pin_memory + non_blocking copies + prefetch gives the reference (pretrain.py:173-179)."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._pc_stage = torch.empty_like(self.pc_in)
            self._img_stage = torch.empty_like(self.img_in)
            self._stage_ready = torch.cuda.Event()
            self._stage_free = torch.cuda.Event()
            self._stage_free.record(torch.cuda.current_stream())
        b = self.b
        cs = self._copy_stream
        cs.wait_event(self._stage_free)                   # the previous staged batch has been consumed
        with torch.cuda.stream(cs):
            self._pc_stage[:b].copy_(pc_t1, non_blocking=True)
            self._pc_stage[b:].copy_(pc_t2, non_blocking=True)
            self._img_stage.copy_(imgs, non_blocking=True)
            self._stage_ready.record(cs)
        self._staged = True

    def step_host_prefetched(self, next_batch=None, lag_losses=False):
        """One step on the batch staged by prefetch_host(); starts the copy of `next_batch` (a (pc_t1, pc_t2, imgs)
        tuple of pinned host tensors) before the step is launched, then reads the losses back (synchronises).

        lag_losses=True: the device-to-host copy of this step's losses is only enqueued; the call returns the losses
        of the PREVIOUS step (None on the first call) once their copy has landed, so the launching thread can run one
        step ahead of the GPU -- what a training loop that logs `loss.item()` of the previous iteration does.
        drain_losses() returns the last step's."""
        if not self._staged:
            raise RuntimeError("step_host_prefetched: no staged batch; call prefetch_host() first")
        cur = torch.cuda.current_stream()
        cur.wait_event(self._stage_ready)
        self.pc_in.copy_(self._pc_stage, non_blocking=True)      # device-to-device, ~25 us
        self.img_in.copy_(self._img_stage, non_blocking=True)
        self._stage_free.record(cur)
        self._staged = False
        if next_batch is not None:
            self.prefetch_host(*next_batch)
        self.step()
        if not lag_losses:
            self.losses_host.copy_(self.losses, non_blocking=True)
            cur.synchronize()
            return self.losses_host
        if self._loss_ring is None:
            self._loss_ring = [(torch.zeros(3, dtype=F32).pin_memory(), torch.cuda.Event()) for _ in range(2)]
        buf, ev = self._loss_ring[self.steps_done & 1]
        buf.copy_(self.losses, non_blocking=True)
        ev.record(cur)
        prev, self._loss_pending = self._loss_pending, (buf, ev)
        if prev is None:
            return None
        prev[1].synchronize()
        return prev[0].clone()            # the ring slot is reused two steps later

    def drain_losses(self):
        """Losses of the last lag_losses step (waits for their device-to-host copy)."""
        prev, self._loss_pending = self._loss_pending, None
        if prev is None:
            return None
        prev[1].synchronize()
        return prev[0].clone()

    def step_host(self, pc_t1, pc_t2, imgs):
        """End-to-end step from (pinned) HOST tensors: H2D copies, the step, D2H read of the losses (synchronises).
        pc_t1/pc_t2 [b,N,3] fp32, imgs [b,3,H,W] fp32 (what the reference's DataLoader yields, pretrain.py:173-179)."""
        b = self.b
        self.pc_in[:b].copy_(pc_t1, non_blocking=True)
        self.pc_in[b:].copy_(pc_t2, non_blocking=True)
        self.img_in.copy_(imgs, non_blocking=True)
        self.step()
        self.losses_host.copy_(self.losses, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.losses_host

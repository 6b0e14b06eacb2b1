"""Whole-step CUDA graphs for models built on the fused PointNet++ path.

A backbone forward+backward is ~150 short kernels on two streams; launched from Python the host needs longer to
enqueue them than the GPU needs to run them (bench.py --no-graph: 9.3 vs 3.7 ms per scene).  `GraphedTrainStep`
captures ONE CUDA graph that holds   forward -> loss -> backward   and replays it per step; with a process group
the replay is followed by ONE NCCL all-reduce (ReduceOp.AVG) of the whole gradient arena.  It replaces, for the part of
train.py that drives the path (`end_points = model(inputs)` ... `loss.backward()` under DistributedDataParallel,
train.py:382,486-520):

  * torch.cuda.make_graphed_callables + DistributedDataParallel (round 1): there every step paid DDP's host-side
    bucket bookkeeping and hooks around two graph replays (SCALE_r01: +0.4 ms flat at N >= 2);
  * DDP's gradient buckets: parameter gradients are *produced in place* in one flat arena (`grad_slot`): the fused
    backward's weight-gradient kernels write straight into their slice, AccumulateGrad adopts the slice as
    `param.grad`, and the arena is reduced in one call -- no per-parameter copies into buckets, no hooks.
    (Capturing the all-reduce INSIDE the graph, issued from gradient hooks on a forked stream, was tried and hung at
    N = 2 on the first attempt; it is not in the tree.)

Semantics kept: after `step(inputs)` every `param.grad` holds the rank-averaged gradient (DDP's contract,
`broadcast_buffers=False` as in train.py:382: BatchNorm running statistics are per rank), the returned loss is
this rank's.  Inputs are copied into static buffers (the copy is the first node of the graph), so callers may pass
fresh tensors every step.

    step = GraphedTrainStep(model, loss_fn, sample_inputs, process_group=None)
    loss = step(cloud)              # tensors in, 0-dim loss tensor out (static buffer: read it before the next step)
    optimizer.step()                # param.grad tensors are stable views into the arena

Geometry prefetch: the first set-abstraction level's FPS + ball query depend on
the cloud alone and are the one part of the step nothing else can hide (0.9 ms of 3.7 at batch 1: a 2047-round
dependent arg-max chain on one 16-SM cluster).  A training loop knows its next batch (the reference prefetches with
DataLoader workers, train.py:257-262), so

    step = GraphedTrainStep(model, loss_fn, sample, prefetch=(model.backbone.sa1, lambda inputs: inputs[0][..., :3]))
    loss = step(cloud_t, next_inputs=(cloud_t1,))      # ... and the following call must be step(cloud_t1, ...)

copies the NEXT batch (host or device) and runs its level-1 geometry as a second small graph on a side stream
underneath the current step's MLPs and backward; the next call finds indices and centres ready.  Every step still
computes its own geometry exactly once, with the same kernels -- only the schedule changes, results are identical.

`GraphedForward` is the inference counterpart (teacher / EMA forward under no_grad, train.py:490-491, and eval).
"""
import torch
import torch.distributed as dist


def grad_slot(param):
    """The arena slice a fused backward should write this parameter's gradient into (None outside a graphed step)."""
    return getattr(param, "_pn2_grad_slot", None)


class _Arena:
    """One flat fp32 buffer holding every trainable parameter's gradient (16-byte aligned slices)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        order = list(reversed(self.params))
        offs, total = [], 0
        for p in order:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4  # 16-byte aligned slices
        dev = order[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.slots = {}
        for p, off in zip(order, offs):
            self.slots[p] = self.flat[off:off + p.numel()].view_as(p)

    def attach(self):
        for p, s in self.slots.items():
            p._pn2_grad_slot = s

    def detach(self):
        for p in self.slots:
            if hasattr(p, "_pn2_grad_slot"):
                del p._pn2_grad_slot


class GraphedTrainStep:
    def __init__(self, model, loss_fn, sample_inputs, process_group=None, warmup=3, prefetch=None):
        """model: nn.Module; loss_fn(model_output) -> scalar tensor; sample_inputs: tuple of CUDA tensors with the
        shapes every later call will use; process_group: None (single rank) or a NCCL group to average gradients over;
        prefetch: None or (sa_module, xyz_of) -- the set-abstraction module whose geometry is computed one step ahead and
        a function mapping the input tuple to its (B, N, 3) coordinates."""
        if not isinstance(sample_inputs, (tuple, list)):
            sample_inputs = (sample_inputs,)
        self.model, self.loss_fn = model, loss_fn
        self.group = process_group
        self.world = dist.get_world_size(process_group) if process_group is not None else 1
        self.static_in = tuple(t.detach().clone().requires_grad_(t.requires_grad) for t in sample_inputs)
        self.arena = _Arena(list(model.parameters()))
        dev = self.static_in[0].device
        self.launches_per_step = None
        self.prefetch = prefetch
        if prefetch is not None:
            import fused
            sa, xyz_of = prefetch
            self.static_next = tuple(t.detach().clone() for t in sample_inputs)  # the batch the geometry graph works on
            self.geom_stream = torch.cuda.Stream(device=dev)
            self.geom_done, self.consumed = torch.cuda.Event(), torch.cuda.Event()
            self.next_ready = False
            warm = torch.cuda.Stream(device=dev)
            warm.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(warm):
                fused.sa_geometry(sa, xyz_of(self.static_next))
            torch.cuda.current_stream(dev).wait_stream(warm)
            torch.cuda.synchronize(dev)
            self.geom_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.geom_graph):
                self.geom_next = fused.sa_geometry(sa, xyz_of(self.static_next))
            self.geom_graph.replay()
            torch.cuda.synchronize(dev)
            self.geom_cur = tuple(t.clone() for t in self.geom_next)  # what the main graph reads
            for dst, src in zip(self.static_in, self.static_next):
                dst.detach().copy_(src)
            sa._pn2_geom = self.geom_cur

        self.arena.attach()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):  # warm-up outside the capture: allocator pools, lazy module state
            for _ in range(warmup):
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        import _pn2
        before = _pn2.kernel_launches()
        with torch.cuda.graph(self.graph):
            self.static_loss = self._run()
        self.launches_per_step = _pn2.kernel_launches() - before
        self.arena.detach()  # the graph holds the pointers; eager calls of the model afterwards behave normally
        if prefetch is not None:
            del prefetch[0]._pn2_geom
            self.launches_per_step += 2  # FPS + ball query of the prefetched level run in the geometry graph

    # one forward + backward; gradients end up in the arena
    def _run(self):
        for p in self.arena.params:
            p.grad = None
        for t in self.static_in:
            t.grad = None
        out = self.model(*self.static_in)
        loss = self.loss_fn(out)
        loss.backward()
        for p in self.arena.params:
            slot = self.arena.slots[p]
            if p.grad is None:
                slot.zero_()  # no gradient this step (unused parameter): contributes zero to the average
            elif p.grad.data_ptr() != slot.data_ptr():
                slot.copy_(p.grad)  # a gradient produced outside the fused path: move it into the arena
            p.grad = slot
        return loss.detach()

    def __call__(self, *inputs, next_inputs=None):
        """One training step on `inputs`.  With prefetch: pass the FOLLOWING batch as `next_inputs` (device or pinned
        host tensors); the next call must then be made with exactly that batch."""
        main = torch.cuda.current_stream(self.static_in[0].device)
        with torch.no_grad():
            if self.prefetch is None:
                for dst, src in zip(self.static_in, inputs):
                    dst.copy_(src, non_blocking=True)
            else:
                if self.next_ready:
                    main.wait_event(self.geom_done)  # batch + geometry staged under the previous step (or a stale prefetch)
                staged = self.next_ready and all(a.data_ptr() == b.data_ptr() and a.shape == b.shape
                                                 for a, b in zip(inputs, self._next_ref))
                if not staged:  # first call, or the caller did not continue with the batch it announced
                    for dst, src in zip(self.static_next, inputs):
                        dst.copy_(src, non_blocking=True)
                    self.geom_graph.replay()
                torch._foreach_copy_([t.detach() for t in self.static_in], list(self.static_next))
                torch._foreach_copy_(list(self.geom_cur), list(self.geom_next))
                self.consumed.record(main)
                self.next_ready = False
        self.graph.replay()
        if self.prefetch is not None and next_inputs is not None:
            with torch.no_grad(), torch.cuda.stream(self.geom_stream):
                self.geom_stream.wait_event(self.consumed)  # the staging buffers were copied out
                for dst, src in zip(self.static_next, next_inputs):
                    dst.copy_(src, non_blocking=True)
                self.geom_graph.replay()
                self.geom_done.record(self.geom_stream)
            self.next_ready, self._next_ref = True, tuple(next_inputs)
        if self.world > 1:  # DDP's contract: every rank ends up with the average gradient
            dist.all_reduce(self.arena.flat, op=dist.ReduceOp.AVG, group=self.group)
        for p, slot in self.arena.slots.items():  # an eager step in between may have re-pointed .grad
            p.grad = slot
        return self.static_loss

    @property
    def input_grads(self):
        """Gradients of the (static) inputs that require grad, valid after a call."""
        return tuple(t.grad for t in self.static_in)


class GraphedForward:
    """no_grad forward of `model` as one CUDA graph (eval / teacher forward).  Outputs: whatever tensors
    `select(model_output)` returns (a tensor or a tuple of tensors), static across replays."""

    def __init__(self, model, sample_inputs, select=lambda out: out, warmup=2):
        if not isinstance(sample_inputs, (tuple, list)):
            sample_inputs = (sample_inputs,)
        self.static_in = tuple(t.clone() for t in sample_inputs)
        dev = self.static_in[0].device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                select(model(*self.static_in))
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = select(model(*self.static_in))

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out

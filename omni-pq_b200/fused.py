"""Fused set-abstraction / feature-propagation execution on libpn2_b200.so.

Replaces, for `PointnetSAModuleVotes` (max pooling) and `PointnetFPModule`, the chain the reference runs
as ~20 separate kernels per layer (pointnet2_modules.py:233-267 / :393-416 -> pointnet2_utils.py:334-359
-> pytorch_utils.py SharedMLP -> cuDNN/ATen), including its autograd backward:

    SA forward :  FPS(+centre gather)  ->  ball query  ->  [gather+centre-subtract fused into GEMM 0]
                  -> BN stats -> [BN+ReLU fused into GEMM l's operand load] ... -> BN+ReLU+max-pool
    FP forward :  three_nn + weights + interpolate (one gather-MAC kernel) -> same MLP chain
    backward   :  pool/ReLU/BN backward folded into the operand loads of the dgrad / wgrad GEMMs,
                  scatter-add into the gathered inputs in the last dgrad's epilogue.

Activations live position-major ([rows, channels], channels padded to a multiple of 4); the API tensors
keep the reference's (B,C,N) layout, and a module output remembers its position-major twin
(`tensor._pn2_pm`) so the next module does not transpose it back.

torch is used for memory, streams, autograd bookkeeping and (SyncBatchNorm only) the cross-rank
all-reduce of the BatchNorm sums; every computation is a libpn2_b200 kernel.
"""
import os

import torch
import torch.distributed as dist
import torch.nn as nn

import _pn2 as K

pad4 = K.pad4

# Geometry (FPS -> ball query of every level) depends on coordinates only.  It is enqueued on a side stream
# so that the FPS chain of level l+1 runs underneath the shared MLP of level l instead of in front of it;
# the MLP stream waits on one event per level.  PN2_GEOM_STREAM=0 keeps everything on the current stream.
_SIDE_STREAM = os.environ.get("PN2_GEOM_STREAM", "1") != "0"
_SIDE_IN_GRAPH = os.environ.get("PN2_GEOM_STREAM_IN_GRAPH", "1") != "0"  # fork/join the side stream inside CUDA-graph capture
_geom_streams = {}


def _geom_stream(device):
    s = _geom_streams.get(device.index)
    if s is None:
        s = _geom_streams[device.index] = torch.cuda.Stream(device=device)
    return s


# ---- what the fused path accepts ------------------------------------------------------------------------
def _mlp_layers(shared_mlp):
    """[(conv, bn)] if every unit is Conv2d 1x1 (no bias) -> BatchNorm -> ReLU, else None."""
    layers = []
    for unit in shared_mlp.children():
        mods = list(unit.children())
        if len(mods) != 3:
            return None
        conv, bnw, act = mods
        if not isinstance(conv, nn.Conv2d) or conv.bias is not None or conv.kernel_size != (1, 1):
            return None
        if conv.stride != (1, 1) or conv.padding != (0, 0) or conv.groups != 1:
            return None
        bn = list(bnw.children())[0] if isinstance(bnw, nn.Sequential) and len(list(bnw.children())) == 1 else bnw
        if not isinstance(bn, nn.modules.batchnorm._BatchNorm) or not bn.affine:
            return None
        if not bn.training and bn.running_mean is None:
            return None
        if not isinstance(act, nn.ReLU):
            return None
        if not (conv.weight.is_cuda and bn.weight.is_cuda and (bn.running_mean is None or bn.running_mean.is_cuda)):
            return None  # parameters / buffers must live on the GPU like the inputs
        layers.append((conv, bn))
    return layers or None


def _ok_tensor(t):
    return t is not None and t.is_cuda and t.dtype == torch.float32


def sa_supported(module, xyz, features):
    if module.npoint is None or module.pooling != 'max' or module.ret_unique_cnt:
        return False
    if getattr(module.grouper, "sample_uniformly", False):
        return False
    if not _ok_tensor(xyz) or (features is not None and not _ok_tensor(features)):
        return False
    if features is None and not module.use_xyz:
        return False
    if not (0 < module.nsample <= 256):
        return False
    return _mlp_layers(module.mlp_module) is not None


def fp_supported(module, unknown, known, unknow_feats, known_feats):
    if known is None or not _ok_tensor(unknown) or not _ok_tensor(known) or not _ok_tensor(known_feats):
        return False
    if unknow_feats is not None and (not _ok_tensor(unknow_feats) or known_feats.shape[1] % 4 != 0):
        return False
    return _mlp_layers(module.mlp) is not None


# ---- helpers -------------------------------------------------------------------------------------------
def _cached_pm(features):
    """Position-major twin remembered by the module that produced `features` (None if absent / stale)."""
    cache = getattr(features, "_pn2_pm", None) if features is not None else None
    if cache is not None and cache[1] == features._version and cache[2] == tuple(features.shape):
        return cache[0]
    return None


def _point_major(features, cached):
    return cached if cached is not None else K.to_point_major(features.detach().contiguous())


def _remember_pm(out, out_pm):
    out._pn2_pm = (out_pm.detach(), out._version, tuple(out.shape))


def _sync_group(bn):
    """Process group to all-reduce BatchNorm sums over, or None (plain BatchNorm / single rank)."""
    if not isinstance(bn, nn.SyncBatchNorm) or not bn.training:
        return None
    if not (dist.is_available() and dist.is_initialized()):
        return None
    group = bn.process_group if bn.process_group is not None else dist.group.WORLD
    return group if dist.get_world_size(group) > 1 else None


def all_reduce_stats(sums, group):
    """SyncBatchNorm exchange: `sums` = this rank's fp64 totals [2C+1] = (sum[C], second sum[C], row count) -- in the
    backward (sum dz[C], sum dz*y[C], unused).  ONE in-place all-reduce per BatchNorm layer and direction (torch's
    SyncBatchNorm issues an all_gather of 2C+1 floats in the forward and an all_reduce of 2C in the backward,
    SURVEY.md 2.5); the reduced row count stays on the device and is consumed by the finalize kernels, so there is no
    host synchronisation and the exchange can be captured in a CUDA graph."""
    dist.all_reduce(sums, group=group)
    return sums


class _PeerExchange:
    """Per process group: this rank's exchange buffer in symmetric memory (mapped into every peer) and what
    pn2_bn_sync_exchange needs to find the peers'.  Built lazily by the first SyncBatchNorm layer that runs; if symmetric
    memory cannot be set up (no peer access, older driver) `ok` stays False and the NCCL all-reduce is used."""
    SLOT = 2 * 2048 + 2  # doubles per slot: up to 2048 channels

    def __init__(self, group, device):
        self.ok, self.why = False, ""
        try:
            import torch.distributed._symmetric_memory as symm
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
            try:
                symm.enable_symm_mem_for_group(group.group_name)  # needed by older torch, a deprecated no-op later
            except Exception:
                pass
            n = 2 * self.SLOT + (self.world + 1) // 2 + 1
            self.buf = symm.empty(n, dtype=torch.float64, device=device)
            self.buf.zero_()
            self.hdl = symm.rendezvous(self.buf, group)
            self.peers = torch.tensor([int(p) for p in self.hdl.buffer_ptrs], dtype=torch.int64, device=device)
            self.ticket = torch.zeros(1, dtype=torch.int32, device=device)
            self.epoch = 0
            torch.cuda.synchronize(device)
            dist.barrier(group)  # every rank's buffer is zeroed before anybody signals into it
            self.ok = True
        except Exception as ex:  # fall back to NCCL; the reason is kept for diagnostics
            self.why = f"{type(ex).__name__}: {ex}"

    def exchange(self, stats, tiles, c, np_, count):
        self.epoch += 1
        return K.bn_sync_exchange(stats, tiles, c, np_, count, self.peers, self.rank, self.world, self.epoch, self.SLOT,
                                  self.ticket)


_peer_exchanges = {}
_PEER_EXCHANGE = os.environ.get("PN2_SYNCBN_PEER", "1") != "0"  # PN2_SYNCBN_PEER=0: always exchange with NCCL
peer_exchange_calls = 0  # exchanges served by the fused kernel (tests read this)


def exchange_stats(stats, tiles, c, np_, count, group):
    """SyncBatchNorm totals over all ranks, [2c+1] fp64 on the device: the fused peer-memory kernel when symmetric memory
    is available and no CUDA graph is being captured (its epoch is a launch argument), else reduce + NCCL all-reduce."""
    global peer_exchange_calls
    if _PEER_EXCHANGE and c <= 2048 and not torch.cuda.is_current_stream_capturing():
        key = (id(group), stats.device.index)
        ex = _peer_exchanges.get(key)
        if ex is None:
            ex = _peer_exchanges[key] = _PeerExchange(group, stats.device)
        if ex.ok:
            peer_exchange_calls += 1
            return ex.exchange(stats, tiles, c, np_, count)
    return all_reduce_stats(K.bn_reduce_stats(stats, tiles, c, np_, count), group)


class _Layer:
    """Per-layer forward state kept for the backward pass."""
    __slots__ = ("cout", "cin", "kp", "np", "xyz_first", "feat_pad", "wt", "wp", "y", "scale", "shift", "mean",
                 "invstd", "training", "count", "count_dev", "group", "gamma", "params")


def _bn_forward(L, bn, stats, tiles, rows):
    """BatchNorm statistics of L.y -> folded scale/shift (running stats in eval mode)."""
    training = bn.training
    L.training, L.group, L.gamma, L.count_dev = training, None, bn.weight, None
    sums = None
    if training:
        pg = _sync_group(bn)
        if pg is not None:
            sums = exchange_stats(stats, tiles, L.cout, L.np, rows, pg)
            L.group, L.count_dev = pg, sums[2 * L.cout:]  # global row count, device resident
        elif rows <= 1:
            raise ValueError("Expected more than 1 value per channel when training")
    L.count = float(rows)
    L.scale, L.shift, L.mean, L.invstd = K.bn_finalize(training, tiles, L.cout, L.np, L.count, stats, sums, bn)


def _prep_layers(layers, kp0, xyz_first, feat_pad):
    """Per-layer geometry + prepared weights (independent of the activations, so callers issue this before they
    wait for anything else)."""
    state, kp = [], kp0
    for li, (conv, bn) in enumerate(layers):
        L = _Layer()
        L.cout, L.cin = conv.weight.shape[0], conv.weight.shape[1]
        L.params = (conv.weight, bn.weight, bn.bias)
        L.kp, L.np = kp, pad4(L.cout)
        L.xyz_first, L.feat_pad = (xyz_first, feat_pad) if li == 0 else (0, 0)
        L.wt, L.wp = K.mlp_prep_weights(conv.weight.detach().view(L.cout, L.cin), L.xyz_first, L.feat_pad, L.kp, L.np)
        state.append(L)
        kp = L.np
    return state


def _run_mlp(layers, state, rows0, nrows):
    """Forward through the conv+BN(+ReLU) stack; fills the per-layer state list made by _prep_layers."""
    assert rows0.cols == state[0].kp
    src = rows0
    for L, (conv, bn) in zip(state, layers):
        L.y, stats, tiles = K.mlp_forward(src, L.kp, L.np, L.wt, L.wp, want_stats=bn.training)
        _bn_forward(L, bn, stats, tiles, nrows)
        src = K.rows_bnrelu(L.y, nrows, L.np, L.np, L.scale, L.shift)
    return state


def _bn_backward(L, stats, tiles):
    """-> (ca, cb, cc, dgamma, dbeta) for dy = ca*dz + cb + cc*y.  Under SyncBatchNorm the sums of (dz, dz*y) are
    all-reduced for the input-gradient coefficients; dgamma / dbeta stay rank-local (DDP averages them), exactly
    like torch.nn.SyncBatchNorm's backward."""
    sums = None
    if L.group is not None:
        sums = exchange_stats(stats, tiles, L.cout, L.np, 0, L.group)
    return K.bn_bwd_finalize(L.training, tiles, L.cout, L.np, L.count, stats, sums, L.gamma.detach(), L.mean, L.invstd,
                             count_dev=L.count_dev, dgamma_out=_slot(L.params[1]), dbeta_out=_slot(L.params[2]))


def _slot(param):
    """Fresh view of the gradient-arena slice of `param` inside a graphed training step (graphed.py), else None.
    A fresh view (not the arena's own tensor object) so that autograd's AccumulateGrad adopts it as param.grad
    instead of cloning it."""
    s = getattr(param, "_pn2_grad_slot", None)
    return None if s is None else s.view(s.shape)


# The weight gradient of a layer and its data gradient both start from dY and do not depend on each other.  On the deep
# levels either GEMM is a few dozen tiles whose time is launch + prologue + tail, and on the shallow ones the tail of one
# fills behind the other, so the weight gradients run on a second stream (forked / joined inside CUDA-graph capture like
# the geometry stream) underneath the data-gradient chain.  Measured per step: never 2.521 ms, <= 8192 positions 2.466,
# always 2.433 (profiles/r2_bench_wgrad_stream_*.json).  PN2_WGRAD_STREAM_ROWS: largest position count that forks.
_WGRAD_STREAM_ROWS = int(os.environ.get("PN2_WGRAD_STREAM_ROWS", str(1 << 30)))
_wgrad_streams = {}


def _wgrad_stream(device):
    s = _wgrad_streams.get(device.index)
    if s is None:
        s = _wgrad_streams[device.index] = torch.cuda.Stream(device=device)
    return s


def _mlp_backward(state, rows0, nrows, gz, out_pm, arg, group, first_dgrad):
    """Backward through the stack.  gz: position-major grad of the pooled output [groups, np_last]
    (overwritten).  first_dgrad(dy, L0) handles the input gradient of layer 0.  Returns the parameter
    gradients [(dW, dgamma, dbeta)] in layer order."""
    last = state[-1]
    groups = nrows // group
    dev = last.y.device
    main = torch.cuda.current_stream(dev)
    side = _wgrad_stream(dev) if 0 < nrows <= _WGRAD_STREAM_ROWS else None
    keep = []  # what the side stream reads stays referenced until the join (its blocks belong to the main stream's pool)
    stats, tiles = K.pool_bwd_prep(gz, out_pm, arg, last.y, groups, group, last.cout, last.np)
    ca, cb, cc, dgamma, dbeta = _bn_backward(last, stats, tiles)
    dy = K.rows_dy(last.y, gz, nrows, last.np, last.np, ca, cb, cc, arg=arg, group=group)
    grads = [None] * len(state)
    for li in range(len(state) - 1, -1, -1):
        L = state[li]
        if li > 0:
            P = state[li - 1]
            a_src = K.rows_bnrelu(P.y, nrows, P.np, P.np, P.scale, P.shift)
        else:
            a_src = rows0
        if side is not None:
            ready = torch.cuda.Event()
            ready.record(main)
            side.wait_event(ready)
            with torch.cuda.stream(side):
                dw = K.mlp_wgrad(dy, a_src, L.cout, L.cin, L.xyz_first, L.feat_pad, L.y.device, out=_slot(L.params[0]))
            dw.record_stream(main)
            keep.append((dy, a_src))
        else:
            dw = K.mlp_wgrad(dy, a_src, L.cout, L.cin, L.xyz_first, L.feat_pad, L.y.device, out=_slot(L.params[0]))
        grads[li] = (dw.view(L.cout, L.cin, 1, 1), dgamma, dbeta)
        if li > 0:
            dz, stats, tiles = K.mlp_dgrad_mask(dy, P.np, L.wp, P.y, P.scale, P.shift, wt=L.wt)
            ca, cb, cc, dgamma, dbeta = _bn_backward(P, stats, tiles)
            dy = K.rows_dy(P.y, dz, nrows, P.np, P.np, ca, cb, cc)
        else:
            first_dgrad(dy, L)
    if side is not None:
        main.wait_stream(side)
    del keep
    return grads


def _flat_params(layers):
    out = []
    for conv, bn in layers:
        out += [conv.weight, bn.weight, bn.bias]
    return out


# ---- set abstraction -------------------------------------------------------------------------------------
class _SAFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, layers, feat_cached, xyz_on_side, xyz, features, inds, *params):
        b, n, _ = xyz.shape
        m, ns = module.npoint, module.nsample

        def geometry():
            xyz_c = xyz.detach().contiguous()
            if inds is None:
                inds_, new_xyz = K.furthest_point_sampling(xyz_c, m, return_xyz=True)
            else:
                inds_ = inds.contiguous()
                new_xyz = K.gather_points(xyz_c.transpose(1, 2).contiguous(), inds_).transpose(1, 2).contiguous()
            return xyz_c, inds_, new_xyz, K.ball_query(new_xyz, xyz_c, module.radius, ns)

        use_xyz = bool(module.use_xyz)

        def activation_independent():
            # feature layout change + weight preparation do not need the geometry: on the MLP stream they run
            # underneath FPS / ball query instead of after them
            if features is not None:
                feat_pm_ = _point_major(features, feat_cached)
                ldf_ = feat_pm_.shape[1]
            else:
                feat_pm_, ldf_ = None, 0
            return feat_pm_, ldf_, _prep_layers(layers, ldf_ + (4 if use_xyz else 0), 1 if use_xyz else 0, ldf_)

        pre = getattr(module, "_pn2_geom", None)  # geometry computed ahead of the step (graphed.GraphedTrainStep prefetch)
        if pre is not None and inds is None and tuple(pre[0].shape) == (b, n, 3) and tuple(pre[3].shape) == (b, m, ns):
            # fresh views: the static buffers themselves must not pick up this call's autograd history
            xyz_c, inds, new_xyz, idx = (t.view(t.shape) for t in pre)
            feat_pm, ldf, state = activation_independent()
        elif _SIDE_STREAM and (_SIDE_IN_GRAPH or not torch.cuda.is_current_stream_capturing()):
            main, side = torch.cuda.current_stream(xyz.device), _geom_stream(xyz.device)
            if not xyz_on_side or inds is not None:
                side.wait_stream(main)  # coordinates (or given indices) were produced on the MLP stream
            with torch.cuda.stream(side):
                xyz_c, inds, new_xyz, idx = geometry()
                ready = torch.cuda.Event()
                ready.record(side)
            xyz.record_stream(side)
            feat_pm, ldf, state = activation_independent()
            main.wait_event(ready)
            for t in (xyz_c, inds, new_xyz, idx):  # allocated on (or, xyz_c, possibly aliasing memory of) the geometry
                t.record_stream(main)              # stream's pool, read by GEMM 0 and the backward on this stream
        else:
            xyz_c, inds, new_xyz, idx = geometry()
            feat_pm, ldf, state = activation_independent()
        c_feat = features.shape[1] if features is not None else 0
        inv_scale = module.radius if module.normalize_xyz else 1.0
        rows0 = K.rows_gather(feat_pm, ldf, ldf, idx, xyz_c, new_xyz, n, m, ns, use_xyz, inv_scale)
        nrows = b * m * ns
        _run_mlp(layers, state, rows0, nrows)
        last = state[-1]
        out_pm, arg = K.bn_relu_pool(last.y, b * m, ns, last.cout, last.np, last.scale, last.shift)
        out = K.to_channel_major(out_pm, b, last.cout, m)
        # outputs go through save_for_backward (a plain attribute would tie output -> grad_fn -> ctx -> output
        # into a reference cycle and the activations would only be freed by the cyclic GC)
        rows0.keep = (feat_pm, idx, xyz_c)
        ctx.save_for_backward(new_xyz, inds, out_pm)
        ctx.pn2 = (state, rows0, nrows, arg, (b, n, m, ns, c_feat, ldf))
        ctx.mark_non_differentiable(inds, out_pm)
        ctx.set_materialize_grads(False)
        return new_xyz, out, inds, out_pm

    @staticmethod
    def backward(ctx, g_new_xyz, g_out, _g_inds, _g_pm):
        with torch.cuda.device(ctx.saved_tensors[2].device):  # kernels launch on the current device's stream
            return _SAFunction._backward(ctx, g_new_xyz, g_out)

    @staticmethod
    def _backward(ctx, g_new_xyz, g_out):
        state, rows0, nrows, arg, (b, n, m, ns, c_feat, ldf) = ctx.pn2
        _new_xyz, inds, out_pm = ctx.saved_tensors
        need_xyz, need_feat = ctx.needs_input_grad[4], ctx.needs_input_grad[5]
        dev = out_pm.device
        dxyz = dfeat = None
        grads = [None] * (3 * len(state))
        if g_out is not None:
            last = state[-1]
            gz = K.to_point_major(g_out.contiguous(), ld=last.np)
            dfeat_pm = K._f32(dev, b * n, ldf, zero=True) if (need_feat and c_feat) else None
            dxyz_pm = K._f32(dev, b * n, 3, zero=True) if (need_xyz and rows0.use_xyz) else None

            def first_dgrad(dy, L0):
                if dfeat_pm is not None or dxyz_pm is not None:
                    K.mlp_dgrad_scatter(dy, L0.kp, L0.wp, rows0, dfeat_pm, dxyz_pm, inds, wt=L0.wt)

            per_layer = _mlp_backward(state, rows0, nrows, gz, out_pm, arg, ns, first_dgrad)
            grads = [g for triple in per_layer for g in triple]
            if dfeat_pm is not None:
                dfeat = K.to_channel_major(dfeat_pm, b, c_feat, n)
            if dxyz_pm is not None:
                dxyz = dxyz_pm.view(b, n, 3)
        if need_xyz and g_new_xyz is not None:
            # new_xyz = xyz[inds]: scatter-add its gradient back (GatherOperation.backward in the reference)
            g = K.gather_points_grad(g_new_xyz.transpose(1, 2).contiguous(), inds, n).transpose(1, 2)
            dxyz = g if dxyz is None else dxyz + g
        if need_xyz and dxyz is None:
            dxyz = torch.zeros(b, n, 3, dtype=torch.float32, device=dev)
        return (None, None, None, None, dxyz, dfeat, None, *grads)


def sa_forward(module, xyz, features, inds):
    layers = _mlp_layers(module.mlp_module)
    on_side = getattr(xyz, "_pn2_side", None) == (xyz._version, tuple(xyz.shape))
    with torch.cuda.device(xyz.device):  # the library launches on the CURRENT device's current stream
        new_xyz, out, inds, out_pm = _SAFunction.apply(module, layers, _cached_pm(features), on_side, xyz, features, inds,
                                                       *_flat_params(layers))
    _remember_pm(out, out_pm)
    if _SIDE_STREAM and getattr(module, "_pn2_geom", None) is None:
        new_xyz._pn2_side = (new_xyz._version, tuple(new_xyz.shape))  # produced on the geometry stream
    return new_xyz, out, inds


def sa_geometry(module, xyz):
    """The coordinate-only part of a set-abstraction level -- FPS (+ centre gather) and the ball query -- as
    (xyz contiguous, inds, new_xyz, idx).  It depends on the cloud alone, so a training loop that knows its NEXT batch
    can run it underneath the current step (graphed.GraphedTrainStep(prefetch=...)); the module consumes the result
    through `module._pn2_geom`."""
    with torch.cuda.device(xyz.device):
        xyz_c = xyz.detach().contiguous()
        inds, new_xyz = K.furthest_point_sampling(xyz_c, module.npoint, return_xyz=True)
        return xyz_c, inds, new_xyz, K.ball_query(new_xyz, xyz_c, module.radius, module.nsample)


# ---- feature propagation -----------------------------------------------------------------------------------
class _FPFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, layers, known_cached, unknown, known, unknow_feats, known_feats, *params):
        b, n, _ = unknown.shape
        m = known.shape[1]
        c2 = known_feats.shape[1]
        known_pm = _point_major(known_feats, known_cached)
        ld2 = known_pm.shape[1]
        c1 = unknow_feats.shape[1] if unknow_feats is not None else 0
        ld1 = pad4(c1)
        ldx = ld2 + ld1
        x = K._f32(unknown.device, b * n, ldx)
        idx, weight = K.fp_interpolate(unknown.detach().contiguous(), known.detach().contiguous(), known_pm, ld2, x, ldx)
        if c1:
            K.to_point_major(unknow_feats.detach().contiguous(), ld=ld1, out=x, col0=ld2)
        nrows = b * n
        rows0 = K.rows_plain(x, nrows, ldx, ldx)
        state = _run_mlp(layers, _prep_layers(layers, ldx, 0, 0), rows0, nrows)
        last = state[-1]
        out_pm, _ = K.bn_relu_pool(last.y, nrows, 1, last.cout, last.np, last.scale, last.shift, want_arg=False)
        out = K.to_channel_major(out_pm, b, last.cout, n)
        ctx.save_for_backward(out_pm)
        ctx.pn2 = (state, rows0, nrows, idx, weight, (b, n, m, c1, c2, ld1, ld2, ldx))
        ctx.mark_non_differentiable(out_pm)
        ctx.set_materialize_grads(False)
        return out, out_pm

    @staticmethod
    def backward(ctx, g_out, _g_pm):
        with torch.cuda.device(ctx.saved_tensors[0].device):
            return _FPFunction._backward(ctx, g_out)

    @staticmethod
    def _backward(ctx, g_out):
        state, rows0, nrows, idx, weight, (b, n, m, c1, c2, ld1, ld2, ldx) = ctx.pn2
        (out_pm,) = ctx.saved_tensors
        need_unknow, need_known = ctx.needs_input_grad[5], ctx.needs_input_grad[6]
        if g_out is None:
            return (None,) * (7 + 3 * len(state))
        last = state[-1]
        gz = K.to_point_major(g_out.contiguous(), ld=last.np)
        box = {}

        def first_dgrad(dy, L0):
            if need_unknow or need_known:
                box["dx"] = K.mlp_dgrad_store(dy, ldx, L0.wp, wt=L0.wt)

        per_layer = _mlp_backward(state, rows0, nrows, gz, out_pm, None, 1, first_dgrad)
        grads = [g for triple in per_layer for g in triple]
        d_unknow = d_known = None
        if "dx" in box:
            dx = box["dx"]
            if need_known:
                dk_pm = K.fp_interpolate_grad(dx, ldx, idx, weight, b, n, m, ld2, ld2)
                d_known = K.to_channel_major(dk_pm, b, c2, m)
            if need_unknow and c1:
                d_unknow = K.to_channel_major(dx, b, c1, n, col0=ld2)
        return (None, None, None, None, None, d_unknow, d_known, *grads)


def fp_forward(module, unknown, known, unknow_feats, known_feats):
    layers = _mlp_layers(module.mlp)
    with torch.cuda.device(unknown.device):
        out, out_pm = _FPFunction.apply(module, layers, _cached_pm(known_feats), unknown, known, unknow_feats, known_feats,
                                        *_flat_params(layers))
    _remember_pm(out, out_pm)
    return out

"""Loss, regulariser, labels and optimiser step of ``learning/runModel.py`` on the B200 kernels.

``cell_loss`` is ``Trainer.calcLossAndOA``'s cell branch (``runModel.py:163-211``): per-cell
``F.kl_div(log_softmax(z), y).sum(1)`` weighted by the raw cell volume (or sqrt / log1p of it),
normalised by the weight sum.  It is an autograd node, so a reference ``Trainer`` may also keep
its own torch loss on the logits this package returns.
"""
from __future__ import annotations

import torch

from ._lib import call, lib, ptr

_WEIGHT_MODE = {None: 0, "none": 0, "sqrt": 1, "log": 2}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _strided(t: torch.Tensor, col0: int):
    """(tensor kept alive, element stride) for a float32 2-D/1-D device tensor column view."""
    if t.dim() == 1:
        return t, 1
    return t, t.stride(0)


class _KLCellLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, gt, weight, mode, group=None, distributed=False):
        n = logits.shape[0]
        dev = logits.device
        logits = logits.contiguous()
        grid = lib().dgnn_small_grid()
        part = torch.empty((grid, 2), dtype=torch.float64, device=dev)
        w_ptr = ptr(weight) if weight is not None else None
        ws = weight.stride(0) if weight is not None else 0
        call("dgnn_kl_loss_fwd", ptr(logits), ptr(gt), gt.stride(0), w_ptr, ws, mode, n, ptr(part), _stream())
        if distributed:           # partitioned scene: (sum w*l, sum w) over the cells of all ranks
            import torch.distributed as dist
            part = part.sum(dim=0, keepdim=True).contiguous()
            dist.all_reduce(part, group=group)
            grid = 1
        sums = torch.empty(3, dtype=torch.float32, device=dev)
        call("dgnn_kl_loss_finalize", ptr(part), grid, ptr(sums), _stream())
        ctx.save_for_backward(logits, gt, sums)
        ctx.weight, ctx.mode = weight, mode
        ctx.mark_non_differentiable(sums)
        return sums[0].clone(), sums

    @staticmethod
    def backward(ctx, gout, _gs):
        logits, gt, sums = ctx.saved_tensors
        weight = ctx.weight
        d = torch.empty_like(logits)
        gout = gout.contiguous().to(torch.float32)
        call("dgnn_kl_loss_bwd", ptr(logits), ptr(gt), gt.stride(0), ptr(weight) if weight is not None else None,
             weight.stride(0) if weight is not None else 0, ctx.mode, logits.shape[0], ptr(sums), ptr(gout), ptr(d),
             _stream())
        return d, None, None, None, None, None


class _PointCellLoss(torch.autograd.Function):
    """bce / mse on one logit per cell (``runModel.py:181-188``)."""

    @staticmethod
    def forward(ctx, logits, target, weight, mode, kind, group=None, distributed=False):
        n = logits.shape[0]
        dev = logits.device
        z = logits.reshape(-1).contiguous()
        grid = lib().dgnn_small_grid()
        part = torch.empty((grid, 2), dtype=torch.float64, device=dev)
        call("dgnn_point_loss_fwd", ptr(z), ptr(target), target.stride(0), ptr(weight) if weight is not None else None,
             weight.stride(0) if weight is not None else 0, mode, kind, n, ptr(part), _stream())
        if distributed:
            import torch.distributed as dist
            part = part.sum(dim=0, keepdim=True).contiguous()
            dist.all_reduce(part, group=group)
            grid = 1
        sums = torch.empty(3, dtype=torch.float32, device=dev)
        call("dgnn_kl_loss_finalize", ptr(part), grid, ptr(sums), _stream())
        ctx.save_for_backward(z, target, sums)
        ctx.weight, ctx.mode, ctx.kind, ctx.shape = weight, mode, kind, logits.shape
        ctx.mark_non_differentiable(sums)
        return sums[0].clone(), sums

    @staticmethod
    def backward(ctx, gout, _gs):
        z, target, sums = ctx.saved_tensors
        weight = ctx.weight
        d = torch.empty_like(z)
        gout = gout.contiguous().to(torch.float32)
        call("dgnn_point_loss_bwd", ptr(z), ptr(target), target.stride(0), ptr(weight) if weight is not None else None,
             weight.stride(0) if weight is not None else 0, ctx.mode, ctx.kind, z.shape[0], ptr(sums), ptr(gout), ptr(d),
             _stream())
        return d.view(ctx.shape), None, None, None, None, None, None


def cell_loss(logits, batch_gt, batch_x, clf, return_sums=False, group=None, distributed=False):
    """``Trainer.calcLossAndOA`` (kl): ``batch_gt[:, :2]`` targets, ``batch_x[:, 0]`` raw volume.
    ``batch_gt`` / ``batch_x`` may live on the host; they are moved to ``logits.device``.
    ``distributed``: the rows are one rank's share of a partitioned scene; the value returned is the loss over
    the cells of all ranks of ``group`` (numerator and normaliser all-reduced)."""
    kind = clf.training.loss
    if kind not in ("kl", "bce", "mse"):
        raise ValueError("%r is not a valid loss. choose either kl, bce or mse" % (kind,))   # runModel.py:189-191
    dev = logits.device
    gt = batch_gt.to(dev, dtype=torch.float32)
    if gt.stride(1) != 1:
        gt = gt.contiguous()
    if clf.regularization.cell_type:
        w = batch_x.to(dev, dtype=torch.float32)
        w = w[:, 0] if w.dim() == 2 else w
        mode = _WEIGHT_MODE[clf.regularization.cell_norm]
    else:
        w, mode = None, 3
    if kind == "kl":
        loss, sums = _KLCellLoss.apply(logits, gt, w, mode, group, distributed)
    else:      # bce supervises with the graph-cut label gt[:, 3], mse with the inside percentage gt[:, 0]
        target = gt[:, 3] if kind == "bce" else gt[:, 0]
        loss, sums = _PointCellLoss.apply(logits, target, w, mode, 0 if kind == "bce" else 1, group, distributed)
    return (loss, sums) if return_sums else loss


class _EdgeReg(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, ei, edge_weight):
        dev = logits.device
        z = logits.contiguous()
        grid = lib().dgnn_small_grid()
        part = torch.empty((grid, 2), dtype=torch.float64, device=dev)
        call("dgnn_edge_reg_fwd", ptr(z), ptr(ei[0]), ptr(ei[1]), ei.shape[1], ptr(part), _stream())
        ctx.save_for_backward(z, ei)
        ctx.scale = float(edge_weight) / max(int(ei.shape[1]), 1)
        return (part[:, 0].sum() * ctx.scale).to(torch.float32)

    @staticmethod
    def backward(ctx, gout):
        z, ei = ctx.saved_tensors
        cnt = torch.zeros(z.shape[0], dtype=torch.int32, device=z.device)
        d = torch.empty_like(z)
        gout = gout.contiguous().to(torch.float32)
        call("dgnn_edge_reg_bwd", ptr(z), ptr(ei[0]), ptr(ei[1]), ei.shape[1], z.shape[0], ctx.scale, ptr(gout), ptr(cnt),
             ptr(d), _stream())
        return d, None, None


def edge_regularization(logits, edge_index, edge_weight):
    """``Trainer.calcRegularization`` (``runModel.py:109-160``): ``mean_e |p0[src_e] - p0[tgt_e]| * edge_weight`` with
    ``p = softmax(logits)``; differentiable (the reference adds it to the training loss once ``regularization.edge_epoch``
    is reached, ``runModel.py:250-255``).  ``logits`` float32[n, 2] on the device."""
    if logits.dim() != 2 or logits.shape[1] != 2:
        raise ValueError("the regulariser needs two logits per cell (loss 'kl')")
    ei = edge_index.to(logits.device, dtype=torch.int64).contiguous()
    return _EdgeReg.apply(logits, ei, edge_weight)


def calc_regularization(logits_cell, data, clf, num_layers):
    """The reference's call (``runModel.py:109-125``): on a sampled batch the innermost adjacency
    ``data.batch_adjs[num_layers]`` (local ids; needs ``graph.additional_num_hops == 1``) over the first ``size[0]``
    logits, otherwise ``data.edge_index`` over all of them."""
    adjs = getattr(data, "batch_adjs", None)
    if adjs:
        adj = adjs[num_layers]
        ei, size = adj[0], adj[2]
        return edge_regularization(logits_cell[:size[0]], ei, clf.regularization.edge_weight)
    return edge_regularization(logits_cell, data.edge_index, clf.regularization.edge_weight)


def labels(logits):
    """``processing/generate_mesh.py:75``: argmax(log_softmax), ties -> 0; uint8 on the device."""
    out = torch.empty(logits.shape[0], dtype=torch.uint8, device=logits.device)
    call("dgnn_argmax_labels", ptr(logits.contiguous()), logits.shape[0], logits.shape[1], ptr(out), _stream())
    return out


def export_scores(logits):
    """``dataLoader.exportScore`` (``processing/data.py:521-535``): the arrays the reference writes next to the logits,
    ``dict(number_of_cells, sigmoid, logits, softmax)`` as host NumPy arrays (computed in one pass on the device)."""
    z = logits.detach().contiguous().float()
    if z.dim() == 1:
        z = z[:, None]
    sig, soft = torch.empty_like(z), torch.empty_like(z)
    call("dgnn_scores", ptr(z), z.shape[0], z.shape[1], ptr(sig), ptr(soft), _stream())
    return dict(number_of_cells=int(z.shape[0]), sigmoid=sig.cpu().numpy(), logits=z.cpu().numpy(),
                softmax=soft.cpu().numpy())


def interface_facets(labels_finite, nfacets):
    """``processing/generate_mesh.py:94-105``: mask of facets whose two cells' labels differ
    (the infinite cell, -1, forced outside)."""
    nf = nfacets.to(labels_finite.device, dtype=torch.int32).contiguous()
    flag = torch.empty(nf.shape[0], dtype=torch.uint8, device=labels_finite.device)
    call("dgnn_interface_facets", ptr(labels_finite), labels_finite.shape[0], ptr(nf), nf.shape[0], ptr(flag), _stream())
    return flag


class BatchPrefetcher:
    """Host -> device staging of training batches on a copy stream, one batch ahead of the compute stream
    (what ``DataLoader(pin_memory=True)`` + ``non_blocking`` copies give the reference trainer).

        pf = BatchPrefetcher(device)
        pf.put(host_batch)                      # async upload of the first batch
        for nxt in batches[1:] + [None]:
            cur = pf.get()                      # device-resident batch; the compute stream waits for its copy
            if nxt is not None: pf.put(nxt)     # next upload overlaps this step
            logits = model(cur); ...

    Tensors shared inside a batch (the same ``edge_index`` in every ``batch_adjs`` entry) are uploaded once and
    stay shared, so the whole-graph fast path of ``SurfaceNet.forward`` still recognises them."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self._pending = None

    def _upload(self, obj, memo):
        if torch.is_tensor(obj):
            key = id(obj)
            if key not in memo:
                memo[key] = obj.to(self.device, non_blocking=True)
            return memo[key]
        if isinstance(obj, dict):
            return type(obj)({k: self._upload(v, memo) for k, v in obj.items()})
        if isinstance(obj, (list, tuple)):
            return type(obj)(self._upload(v, memo) for v in obj)
        return obj

    def put(self, host_batch):
        with torch.cuda.stream(self.stream):
            memo = {}
            dev_batch = self._upload(host_batch, memo)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._pending = (dev_batch, ev, list(memo.values()))

    def get(self):
        dev_batch, ev, tensors = self._pending
        self._pending = None
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in tensors:                       # allocated on the copy stream, consumed on the compute stream
            if t.is_cuda:
                t.record_stream(cur)
        return dev_batch


class Adam(torch.optim.Optimizer):
    """``torch.optim.Adam(params, lr)`` with the reference's defaults (``runModel.py:290``), all
    parameter tensors updated by ONE kernel launch.  The step count and the hyper-parameters live in device memory
    (``dgnn_adam_multi_dev``) and the pointer table is staged through pinned host memory, so ``step()`` can be captured
    in a CUDA graph (``GraphedStep``) and replayed; ``lr`` may be changed between steps as the reference does
    (``adjust_learning_rate``, ``runModel.py:95-99``)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._step = 0

    def _group_state(self, group, dev):
        gs = group.get("_dev")
        if gs is None:
            gs = group["_dev"] = dict(step=torch.zeros(1, dtype=torch.int64, device=dev),
                                      hyper=torch.empty(4, dtype=torch.float32, device=dev), hyper_host=None,
                                      table=None, table_host=None, rows=None)
        return gs

    def refresh_hyper(self):
        """Upload lr / betas / eps of every group if they changed since the last step (called by ``step`` and, before a
        replay, by ``GraphedStep`` - a captured step never runs this Python again)."""
        for group in self.param_groups:
            gs = group.get("_dev")
            if gs is None:
                continue
            b1, b2 = group["betas"]
            hyper = (float(group["lr"]), float(b1), float(b2), float(group["eps"]))
            if gs["hyper_host"] != hyper:
                if torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("Adam hyper-parameters changed during CUDA graph capture")
                gs["hyper"].copy_(torch.tensor(hyper, dtype=torch.float32))
                gs["hyper_host"] = hyper

    @torch.no_grad()
    def step(self, closure=None):
        self._step += 1
        capturing = torch.cuda.is_current_stream_capturing()
        for group in self.param_groups:
            rows, max_n = [], 0
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                st["_g"] = g  # keep alive until the launch is enqueued
                rows.append([p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                             p.numel()])
                max_n = max(max_n, p.numel())
            if not rows:
                continue
            dev = group["params"][0].device
            gs = self._group_state(group, dev)
            self.refresh_hyper()
            if gs["rows"] != rows:                             # gradient buffers moved (or first step): restage the table
                if gs["table_host"] is None or gs["table_host"].shape[0] != len(rows):
                    gs["table_host"] = torch.empty((len(rows), 5), dtype=torch.int64).pin_memory()
                    gs["table"] = torch.empty((len(rows), 5), dtype=torch.int64, device=dev)
                gs["table_host"].copy_(torch.tensor(rows, dtype=torch.int64))
                # pinned -> device: a memcpy node under capture; blocking otherwise (the pinned buffer is rewritten by the
                # next restaging, which must not overtake an asynchronous copy)
                gs["table"].copy_(gs["table_host"], non_blocking=capturing)
                gs["rows"] = rows
            gs["step"] += 1
            call("dgnn_adam_multi_dev", ptr(gs["table"]), len(rows), max_n, ptr(gs["hyper"]), ptr(gs["step"]), _stream())
            for p in group["params"]:           # the kernel wrote the parameters behind autograd's back: record the
                if p.grad is not None:          # in-place update so that version-keyed caches (packed weights) notice
                    torch.autograd.graph.increment_version(p)


class GraphedStep:
    """One training step (forward, loss, backward, optimiser) of a FIXED device-resident batch captured into a CUDA
    graph and replayed: the ~140 kernel launches of a step become one graph launch (no launch gaps, no Python between
    kernels).  The reference trains for many epochs on one collated batch of all training graphs (``run.py:59-61``), so
    the captured step is replayed as long as the batch object and its layout do not change; a different batch captures
    a new graph.

        step = GraphedStep(lambda: cell_loss(net(batch), batch.all.y, batch.all.x, clf), optimizer)
        for epoch in ...: loss = step()          # a device scalar (updated in place by every replay)
    """

    def __init__(self, loss_fn, optimizer, post_backward=None, warmup=3):
        self.loss_fn, self.opt, self.post_backward, self.warmup = loss_fn, optimizer, post_backward, warmup
        self.graph = None
        self.loss = None
        self.calls = 0

    def _eager(self):
        loss = self.loss_fn()
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        if self.post_backward is not None:
            self.post_backward()
        self.opt.step()
        return loss

    def __call__(self):
        """The first ``warmup`` calls run eagerly (they are ordinary training steps: graph plans, packed weights, optimiser
        state and the allocator settle), the next call captures the step and every call from then on replays it."""
        if self.calls < self.warmup:
            self.calls += 1
            return self._eager()
        if self.graph is None:
            torch.cuda.synchronize()
            self.opt.zero_grad(set_to_none=True)              # gradients are allocated inside the graph's private pool
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):                # records, does not execute
                self.loss = self._eager()
        self.calls += 1
        if hasattr(self.opt, "refresh_hyper"):
            self.opt.refresh_hyper()                          # e.g. the reference's adjust_learning_rate between epochs
        self.graph.replay()
        for group in self.opt.param_groups:                   # the replay rewrote the parameters: version-keyed caches
            for p in group["params"]:                         # (packed weights of a later eval pass) must notice
                torch.autograd.graph.increment_version(p)
        return self.loss

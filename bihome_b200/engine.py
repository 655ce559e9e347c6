"""Config-driven model factory and training step (reference train.py:544-757 main, :284-429 train_one_epoch).

The YAML contract is the reference's: MODEL.BACKBONE / MODEL.HEAD are passed as **kwargs to
``<package>.backbones.<NAME>.Model`` and ``<package>.heads.<NAME>.Model(backbone, **kwargs)``; the model is
``nn.Sequential(backbone, head)`` and the batch is one dict flowing through both.
"""
import importlib

import torch
import yaml


def load_config(path):
    with open(path, 'r') as f:
        return yaml.full_load(f)


def build_model(config, pretrained=None):
    """nn.Sequential(backbone, head) from a reference-style config dict.  ``pretrained=False`` forces random init
    of both the backbone trunk and the frozen extractor (no network on the GPU boxes)."""
    bcfg, hcfg = dict(config['MODEL']['BACKBONE']), dict(config['MODEL']['HEAD'])
    if pretrained is not None:
        bcfg['PRETRAINED_RESNET'] = bool(pretrained) and bcfg.get('PRETRAINED_RESNET', False)
        hcfg['AUXILIARY_RESNET_PRETRAINED'] = bool(pretrained)
    backbone = importlib.import_module('bihome_b200.backbones.{}'.format(bcfg['NAME'])).Model(**bcfg)
    head = importlib.import_module('bihome_b200.heads.{}'.format(hcfg['NAME'])).Model(backbone, **hcfg)
    return torch.nn.Sequential(backbone, head)


def build_optimizer(config, model, capturable=False):
    s = config['SOLVER']
    if s['OPTIMIZER'] != 'Adam':
        raise NotImplementedError('I do not have this solver implemented yet.')
    wd = float(s['L2_WEIGHT_DECAY']) if 'L2_WEIGHT_DECAY' in s else 0
    # every parameter, frozen extractor included, exactly like the reference (train.py:703-707): the param-group layout
    # is part of the checkpoint format ('optimizer' entry of model_%06d.pth); frozen parameters never get a .grad, so
    # Adam skips them and keeps no state for them
    opt = torch.optim.Adam(model.parameters(), lr=s['LR'], betas=(s['MOMENTUM_1'], s['MOMENTUM_2']), weight_decay=wd,
                           capturable=capturable, foreach=True)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=s['MILESTONES'], gamma=s['LR_DECAY'])
    return opt, sched


def data_parallel(model, local_rank=None, bucket_cap_mb=48):
    """One process per GPU: DDP's bucketed NCCL all-reduce of the learnable backbone gradients (42.3 MB for Zeng) over
    NVLink, overlapped with backward.  ``broadcast_buffers=False`` keeps BatchNorm running statistics rank-local, as the
    per-rank batches of the reference's per-GPU BatchNorm are (no SyncBN in the reference) and
    removes the per-forward broadcast of the model's 183 buffers; the frozen extractor and the fixed control flow make
    the graph static; one 48 MB bucket cap turns the Zeng gradients into a single late all-reduce plus the stem's."""
    on_gpu = next(model.parameters()).is_cuda
    return torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank] if on_gpu else None, gradient_as_bucket_view=True,
                                                     broadcast_buffers=False, static_graph=True, bucket_cap_mb=bucket_cap_mb)


def train_step(model, data, optimizer, scheduler=None, gradient_clip=-1):
    """one iteration of the reference's hot loop (train.py:305-387) for the string losses ('biHomE', ...):
    returns (loss, delta_gt, delta_hat) as device tensors -- no host sync."""
    optimizer.zero_grad(set_to_none=True)
    loss, delta_gt, delta_hat = model(data)
    loss.backward()
    if gradient_clip > 0:
        torch.nn.utils.clip_grad_norm_(model.parameters(), gradient_clip)
    optimizer.step()
    if scheduler is not None:
        scheduler.step()
    return loss, delta_gt, delta_hat


def _world():
    import torch.distributed as dist
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def flatten_gradients(params):
    """ONE flat buffer behind the gradients of `params`: every ``p.grad`` becomes a view of it with the parameter's own
    strides (channels-last kernels keep their layout), zero-filled.  Backward then accumulates in place into the views, the
    data-parallel exchange is a single all-reduce of the flat buffer, and a CUDA graph of the step sees static addresses."""
    params = list(params)
    flat = torch.zeros(sum(p.numel() for p in params), device=params[0].device, dtype=params[0].dtype)
    off = 0
    for p in params:
        dense = p.is_contiguous() or (p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last))
        stride = p.stride() if dense else torch.empty(p.shape).stride()
        p.grad = torch.as_strided(flat, p.shape, stride, off)
        off += p.numel()
    return flat


def average_gradients(params, flat=None):
    """data-parallel gradient exchange without a DDP wrapper: mean over the ranks (what DDP computes), as ONE all-reduce when
    the gradients are views of a flat buffer (flatten_gradients), coalesced otherwise.  No-op in a single process."""
    import torch.distributed as dist
    world = _world()
    if world == 1:
        return
    if flat is None:
        grads = [p.grad for p in params if p.grad is not None]
        if not grads:
            return
        flat = torch._utils._flatten_dense_tensors(grads)
        dist.all_reduce(flat)
        flat.div_(world)
        for g, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
            g.copy_(f)
        return
    if dist.get_backend() == 'nccl':
        dist.all_reduce(flat, op=dist.ReduceOp.AVG)
    else:
        dist.all_reduce(flat)
        flat.div_(world)


def sync_module_state(model, src=0):
    """what DistributedDataParallel does once at construction: parameters and buffers of every rank <- rank `src`"""
    import torch.distributed as dist
    if _world() == 1:
        return
    with torch.no_grad():
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src)


class GraphedStep:
    """Forward + backward of ``nn.Sequential(backbone, head)`` captured ONCE in a CUDA graph and replayed per step
    ("CUDA-graph capture of the whole head").  The ~2 000 launches of a step (cuDNN, ATen and the
    C-ABI kernels, which launch on torch's current stream and therefore land in the capture) become one graph launch.
    At B = 256 on a B200 the eager step is bound by the host's launch rate for a tenth of its time (58.3 ms eager against
    53.1 ms replayed, profiles/r04e_*); at small batches the replay is the whole step time.

    The batch lives in static device buffers (``__call__`` copies the new batch into them; ``static`` exposes them so that a
    loader can write there directly).  The parameter gradients are views of ONE flat buffer (flatten_gradients) that the
    graph zeroes and accumulates into -- never set them to None.  The optimizer and the scheduler stay eager, so
    learning-rate schedules and checkpoints work unchanged.  Random draws inside the head (``torch.multinomial``) use
    torch's graph-safe Philox state.

    Data parallel (one process per GPU, torch.distributed initialised): the model is NOT wrapped in DDP -- bucket hooks
    are not capturable this way.  The parameters and buffers are broadcast from rank 0 once, every rank replays its own
    graph on its own shard, and the gradient exchange is one NCCL all-reduce (mean) of the flat buffer right after the
    replay: 42.3 MB for the Zeng backbone, ~0.2 ms on NVLink against a 53 ms step, so not overlapping it with backward
    costs less than DDP's per-bucket hooks did.  BatchNorm statistics stay rank-local, as under
    ``data_parallel(broadcast_buffers=False)``.
    """

    def __init__(self, model, example_batch, warmup=3):
        if not torch.cuda.is_available():
            raise RuntimeError('bihome_b200: GraphedStep needs a CUDA device (no CPU fallback)')
        self.model = model
        self.world = _world()
        if self.world > 1:
            sync_module_state(model)
        self.static = {k: v.clone() for k, v in example_batch.items() if torch.is_tensor(v)}
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.flat = flatten_gradients(self.params)
        # the warm-up passes must not count: BatchNorm running statistics (and their batch counters) are put back afterwards
        buffers = [(b, b.detach().clone()) for b in model.buffers()]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):       # warm-up off the capture: lazy initialisation (cuDNN plans, function attributes)
            for _ in range(warmup):
                self.flat.zero_()
                loss, _, _ = model(dict(self.static))
                loss.backward()
        torch.cuda.current_stream().wait_stream(side)
        from . import cabi
        before = cabi.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.flat.zero_()
            self.loss, self.delta_gt, self.delta_hat = model(dict(self.static))
            self.loss.backward()
        self.launches_per_replay = cabi.launch_count() - before      # C-ABI kernels inside one replay (bench.py's gpu_launches)
        with torch.no_grad():
            for b, saved in buffers:
                b.copy_(saved)

    def replay(self):
        """one forward + backward on whatever the static buffers hold, then the data-parallel gradient mean"""
        self.graph.replay()
        average_gradients(self.params, self.flat)
        return self.loss, self.delta_gt, self.delta_hat

    def __call__(self, batch):
        for k, v in self.static.items():
            v.copy_(batch[k], non_blocking=True)
        return self.replay()


def graphed_train_step(step, data, optimizer, scheduler=None, gradient_clip=-1):
    """train_step with the forward + backward replayed from a GraphedStep"""
    loss, delta_gt, delta_hat = step(data)
    if gradient_clip > 0:
        torch.nn.utils.clip_grad_norm_(step.model.parameters(), gradient_clip)
    optimizer.step()
    if scheduler is not None:
        scheduler.step()
    return loss, delta_gt, delta_hat

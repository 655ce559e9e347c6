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


class GraphedStep:
    """Forward + backward of ``nn.Sequential(backbone, head)`` captured ONCE in a CUDA graph and replayed per step
    ("CUDA-graph capture of the whole head").  The ~2 000 launches of a step (cuDNN, ATen and the
    C-ABI kernels, which launch on torch's current stream and therefore land in the capture) become one graph launch:
    at small batches, where the step is launch-bound, that is the whole step time.

    The batch lives in static device buffers (``__call__`` copies the new batch into them), the parameter gradients in
    the graph's private pool (never set them to None); the optimizer and the scheduler stay eager, so learning-rate
    schedules and checkpoints work unchanged.  Random draws inside the head (``torch.multinomial``) use torch's
    graph-safe Philox state.  Single process only: under DDP the bucketed all-reduce hooks are not capturable this way.
    """

    def __init__(self, model, example_batch, warmup=3):
        if not torch.cuda.is_available():
            raise RuntimeError('bihome_b200: GraphedStep needs a CUDA device (no CPU fallback)')
        self.model = model
        self.static = {k: v.clone() for k, v in example_batch.items() if torch.is_tensor(v)}
        params = [p for p in model.parameters() if p.requires_grad]
        # the warm-up passes must not count: BatchNorm running statistics (and their batch counters) are put back afterwards
        buffers = [(b, b.detach().clone()) for b in model.buffers()]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):       # warm-up off the capture: lazy initialisation (cuDNN plans, function attributes)
            for _ in range(warmup):
                for p in params:
                    p.grad = None
                loss, _, _ = model(dict(self.static))
                loss.backward()
        torch.cuda.current_stream().wait_stream(side)
        for p in params:
            p.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, self.delta_gt, self.delta_hat = model(dict(self.static))
            self.loss.backward()
        with torch.no_grad():
            for b, saved in buffers:
                b.copy_(saved)

    def __call__(self, batch):
        for k, v in self.static.items():
            v.copy_(batch[k], non_blocking=True)
        self.graph.replay()
        return self.loss, self.delta_gt, self.delta_hat


def graphed_train_step(step, data, optimizer, scheduler=None, gradient_clip=-1):
    """train_step with the forward + backward replayed from a GraphedStep"""
    loss, delta_gt, delta_hat = step(data)
    if gradient_clip > 0:
        torch.nn.utils.clip_grad_norm_(step.model.parameters(), gradient_clip)
    optimizer.step()
    if scheduler is not None:
        scheduler.step()
    return loss, delta_gt, delta_hat
